"""Tensor-core (tcgen05 / TMEM) scoring + top-k against its own accumulators (bit-exact
selection) and against the fp32 oracle (2e-2 relative, BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from furusato_recommend_b200 import ops  # noqa: E402
from helpers import gap_rule  # noqa: E402

DEV = "cuda:0"


def _stable_topk(scores, pos_lists, k, mask=-1024.0):
    s = scores.clone()
    for r, items in enumerate(pos_lists):
        if len(items):
            s[r, torch.as_tensor(np.asarray(items), device=s.device, dtype=torch.long)] = mask
    vals, idx = torch.sort(s, dim=1, descending=True, stable=True)
    return idx[:, :k], vals[:, :k]


def _case(U, m, d, seed, quantise=False, dense_pos=False):
    gen = torch.Generator(device="cpu").manual_seed(seed)
    ue = torch.randn(U, d, generator=gen).to(DEV)
    ie = torch.randn(m, d, generator=gen).to(DEV)
    if quantise:  # exact ties survive the bf16 rounding
        ue, ie = (ue * 2).round() / 2, (ie * 2).round() / 2
    rng = np.random.default_rng(seed)
    if dense_pos:
        lists = [np.sort(rng.choice(m, rng.integers(m // 2 + 5, m - 4), replace=False)) for _ in range(U)]
    else:
        lists = [np.unique(rng.integers(0, m, rng.integers(0, 30))) for _ in range(U)]
    rowptr = torch.tensor(np.concatenate([[0], np.cumsum([len(x) for x in lists])]), dtype=torch.int64, device=DEV)
    flat = torch.tensor(np.concatenate(lists), dtype=torch.int32, device=DEV)
    ids = torch.randperm(U, generator=gen).to(DEV)
    return ue, ie, ids, rowptr, flat, lists


@pytest.mark.parametrize("shape", [
    (300, 400, 32, 20, False), (1000, 5000, 64, 20, True), (257, 1111, 128, 20, False),
    (130, 700, 64, 50, False), (64, 40, 64, 20, False), (128, 256, 64, 1, False), (129, 513, 64, 24, True),
    # long sweeps: the candidate path after warm-up (argmax descent, several equal maxima in one chunk, heap
    # folds) and, at d = 128, the early hand-back of the TMEM stage with values picked from registers
    (300, 50000, 64, 20, True), (200, 30000, 128, 20, True), (260, 40000, 128, 5, False),
])
@pytest.mark.parametrize("prec", ["bf16", "f16"])
def test_tensor_core_topk(shape, prec):
    U, m, d, k, quant = shape
    ue, ie, ids, rowptr, flat, lists = _case(U, m, d, seed=U + m + k, quantise=quant, dense_pos=(m == 40))
    idx, val, dense = ops.score_topk(ue, ie, ids, rowptr, flat, k, precision=prec, return_scores=True)
    torch.cuda.synchronize()
    # (1) the accumulators are the product of the bf16-rounded operands (fp32 accumulate, or f16
    #     accumulate: one rounding to 11 bits per K=16 step)
    lo = torch.bfloat16 if prec == "bf16" else torch.float16
    ref = ue[ids].to(lo).float() @ ie.to(lo).float().t()
    err = float((dense - ref).abs().max() / ref.abs().max())
    assert err < (1e-5 if prec == "bf16" else 2e-3), f"accumulator mismatch {err:.3e}"
    # (2) selection is bit-exact on the kernel's own scores, ties to the lowest id
    widx, wval = _stable_topk(dense, [lists[u] for u in ids.cpu().tolist()], k)
    assert torch.equal(idx.long(), widx), f"{int((idx.long() != widx).sum())} ids differ"
    assert torch.equal(val, wval)
    # (3) within the bf16 tolerance of the fp32 scores
    f32 = ue[ids] @ ie.t()
    assert float((dense - f32).abs().max() / f32.abs().max()) < 2e-2
    # (4) the non-debug entry returns the same lists
    idx2, val2 = ops.score_topk(ue, ie, ids, rowptr, flat, k, precision=prec)
    assert torch.equal(idx2, idx) and torch.equal(val2, val)


def test_tensor_core_topk_many_tiles_and_users():
    """Several CTAs, > stages item tiles (smem ring wraps, both TMEM stages reused many times)."""
    ue, ie, ids, rowptr, flat, lists = _case(700, 9000, 64, seed=11)
    idx, val, dense = ops.score_topk(ue, ie, ids, rowptr, flat, 20, precision="bf16", return_scores=True)
    widx, wval = _stable_topk(dense, [lists[u] for u in ids.cpu().tolist()], 20)
    assert torch.equal(idx.long(), widx) and torch.equal(val, wval)
    # against the exact fp32 scores: nothing selected from below the fp32 k-th value by
    # more than the bf16 tolerance, and identical ids wherever the fp32 ranks are separated by more
    f32 = ue[ids] @ ie.t()
    for r, u in enumerate(ids.cpu().tolist()):
        if len(lists[u]):
            f32[r, torch.as_tensor(np.asarray(lists[u]), device=f32.device, dtype=torch.long)] = -1024.0
    tol = 2e-2 * float(f32[f32 > -1000].abs().max())
    bad, compared, mism = gap_rule(idx, f32, 20, tol)
    assert bad == 0 and mism == 0, (bad, compared, mism)
    fidx, fval = ops.score_topk(ue, ie, ids, rowptr, flat, 20, precision="fp32")
    bad, compared, mism = gap_rule(fidx, f32, 20, 1e-5 * float(f32[f32 > -1000].abs().max()))
    assert bad == 0 and mism == 0 and compared > 0.95 * fidx.numel(), (bad, compared, mism)


def test_tensor_core_rejects_oversized_k():
    ue, ie, ids, rowptr, flat, _ = _case(64, 400, 64, seed=3)
    from furusato_recommend_b200 import _lib
    with pytest.raises(_lib.LgcnLibraryError):
        ops.score_topk(ue, ie, ids, rowptr, flat, 120, precision="bf16")
