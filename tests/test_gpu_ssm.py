"""Sampled-softmax kernel (csrc/ssm.cu, BASELINE configs[3]) against the plain-torch statement of
the same objective.  PARITY UNPINNED: the reference's model/lgcnssm.py has no working
sampled-softmax arithmetic (its `softmax_loss` is the BPR loss and OneEpoch raises NameError), so
the torch formula below is our own specification; the reference's actual arithmetic for this file —
BPR over flat triples — is covered by the BPR tests (LightGCNSSM without ssm_true_softmax)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from furusato_recommend_b200 import LightGCNSSM, UniformSample, ops  # noqa: E402
from furusato_recommend_b200.dataloader import BasicDataset  # noqa: E402
from furusato_recommend_b200.synthetic import bipartite  # noqa: E402

DEV = "cuda:0"


def _torch_ssm(out, emb, users, pos, neg, n, J, tau):
    u, p = out[users[::J]], out[n + pos[::J]]
    q = out[n + neg].view(-1, J, out.shape[1])
    logits = torch.cat([(u * p).sum(1, keepdim=True), torch.einsum("bd,bjd->bj", u, q)], dim=1) / tau
    loss = (torch.logsumexp(logits, dim=1) - logits[:, 0]).mean()
    B = u.shape[0]
    reg = 0.5 * (emb[users[::J]].pow(2).sum() + emb[n + pos[::J]].pow(2).sum() + emb[n + neg].pow(2).sum()) / B
    return loss, reg


@pytest.mark.parametrize("d,J,B,tau", [(64, 8, 50, 1.0), (32, 16, 33, 0.5), (128, 256, 20, 1.0), (64, 1, 64, 2.0)])
def test_ssm_kernel_matches_torch_formula(d, J, B, tau):
    gen = torch.Generator().manual_seed(d + J)
    n, m = 70, 90
    out = (torch.randn(n + m, d, generator=gen) * 0.3).to(DEV).requires_grad_(True)
    emb = (torch.randn(n + m, d, generator=gen) * 0.1).to(DEV)
    users = torch.randint(0, n, (B,), generator=gen).repeat_interleave(J).to(DEV)     # duplicates across samples
    pos = torch.randint(0, m, (B,), generator=gen).repeat_interleave(J).to(DEV)
    neg = torch.randint(0, m, (B * J,), generator=gen).to(DEV)
    loss, reg = _torch_ssm(out, emb, users, pos, neg, n, J, tau)
    (gref,) = torch.autograd.grad(loss, out)
    G = torch.zeros(n + m, d, device=DEV)
    cnt = torch.zeros(n + m, dtype=torch.int32, device=DEV)
    lo = torch.zeros(4, device=DEV)
    work = torch.empty(2 * B, device=DEV)
    wc = torch.zeros(2, dtype=torch.int32, device=DEV)
    decay = 1e-3
    ops.ssm_fwd_bwd(out.detach(), emb, users, pos, neg, J, n, tau, decay, G, cnt, lo, work, wc)
    assert abs(float(lo[0]) - float(loss)) < 1e-5 * abs(float(loss))
    assert abs(float(lo[1]) - float(reg)) < 1e-5 * abs(float(reg))
    assert abs(float(lo[2]) - float(loss + decay * reg)) < 1e-5 * abs(float(loss))
    assert float((G - gref).abs().max() / gref.abs().max()) < 1e-5
    want = torch.bincount(torch.cat([users[::J], n + pos[::J], n + neg]), minlength=n + m)
    assert torch.equal(cnt.long(), want)
    assert int(wc[1]) == 0
    # an out-of-range id is skipped, counted and poisons the loss (the reference's IndexError)
    neg2 = neg.clone(); neg2[3] = m
    G.zero_(); cnt.zero_()
    ops.ssm_fwd_bwd(out.detach(), emb, users, pos, neg2, J, n, tau, decay, G, cnt, lo, work, wc)
    assert int(wc[1]) == 1 and bool(torch.isnan(lo[0]))


def test_ssm_fused_step_matches_autograd_step():
    n, m, tu, ti, su, si = bipartite(3000, 2000, 80000, seed=3)
    J, B = 16, 128
    base = dict(recdim=64, layer=3, lr=1e-3, decay=1e-4, bpr_batch_size=B, device=DEV, test_u_batch_size=1000,
                neg_size=J, ssm_true_softmax=True)
    ds = BasicDataset(n, m, tu.numpy(), ti.numpy(), su.numpy(), si.numpy(), config=base, device=DEV)
    S = UniformSample(ds, neg_ratio=J, seed=1, epoch=0, count=B)
    assert S.shape == (B * J, 3)
    users, pos, neg = (S[:, k].contiguous() for k in range(3))
    torch.manual_seed(0)
    fused = LightGCNSSM(base, ds).train()
    ref = LightGCNSSM(dict(base, ssm_autograd=True), ds).train()
    with torch.no_grad():
        ref.all_embedding.weight.copy_(fused.all_embedding.weight)
    for _ in range(2):
        lf = float(fused.stageOne(users, pos, neg))
        with torch.enable_grad():
            lr_ = float(ref.stageOne(users, pos, neg))
        assert abs(lf - lr_) < 1e-5 * abs(lr_), (lf, lr_)
    wf, wr = fused.all_embedding.weight.detach(), ref.all_embedding.weight.detach()
    assert float((wf - wr).abs().max() / wr.abs().max()) < 1e-4
    assert float(((wf - wr).abs() > 0.05 * 1e-3).float().mean()) < 1e-3
    ep = float(fused.OneEpoch(users, pos, neg))       # lgcnssm.py:135-153 batching on the fused kernel
    assert np.isfinite(ep) and ep > 0
