"""CUDA parity of the f-4 variants (rAdjGCN, RGCN, PyG-form LightGCN, capped sampler) against the
vectors frozen from the live reference classes.  fp32 tolerance 1e-5 relative (north_star);
sampled triples bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from furusato_recommend_b200 import RGCN, LightGCN, UniformSampleCapped, rAdjGCN  # noqa: E402
from furusato_recommend_b200.dataloader import BasicDataset  # noqa: E402
from oracle import lgcn_oracle as orc  # noqa: E402
from test_gpu_parity import assert_close, batch, golden_dataset  # noqa: E402

DEV = "cuda:0"


def _cfg(golden, **extra):
    d, K, B = (int(x) for x in golden["config"])
    lr, decay = (float(x) for x in golden["hyper"])
    return dict(recdim=d, layer=K, lr=lr, decay=decay, bpr_batch_size=B, device=DEV, **extra)


def _load(model, golden):
    with torch.no_grad():
        model.all_embedding.weight.copy_(torch.from_numpy(golden["E0"]))
    return model


def _check(model, golden, variants, tag):
    model.eval()
    with torch.no_grad():
        u, i = model.forward()
    assert_close(u, variants[f"{tag}_users"], what=f"{tag} users")
    assert_close(i, variants[f"{tag}_items"], what=f"{tag} items")
    # autograd path: loss, reg and the gradient of all_embedding (transpose propagation)
    model.train()
    bu, bp, bn = batch(golden)
    loss, reg = model.bpr_loss(bu, bp, bn)
    assert abs(loss.item() - float(variants[f"{tag}_loss"])) <= 1e-5 * abs(float(variants[f"{tag}_loss"]))
    assert abs(reg.item() - float(variants[f"{tag}_reg"])) <= 1e-5 * abs(float(variants[f"{tag}_reg"]))
    model.optim.zero_grad()
    (loss + float(model.config["decay"]) * reg).backward()
    assert_close(model.all_embedding.weight.grad, variants[f"{tag}_grad"], what=f"{tag} grad")
    model.all_embedding.weight.grad = None
    # fused path: two stageOne steps (Adam in the last backward epilogue)
    for step in (1, 2):
        l = model.stageOne(bu, bp, bn)
        ref = float(variants[f"{tag}_step{step}_loss"])
        assert abs(l.item() - ref) <= 1e-5 * abs(ref)
        E = variants[f"{tag}_E{step}"]
        assert float((model.all_embedding.weight.detach().cpu() - torch.from_numpy(E)).abs().max()) < 2e-6


@pytest.mark.parametrize("cuda_graph", [False, True])
def test_radj_matches_live_reference(golden, variants, cuda_graph):
    model = _load(rAdjGCN(_cfg(golden, r=float(variants["r"]), cuda_graph=cuda_graph), golden_dataset(golden)), golden)
    _check(model, golden, variants, "radj")


def test_radj_half_is_lightgcn(golden):
    """r = 0.5 is the symmetric normalisation (model/radj.py:32-36 with r = 1-r)."""
    a = _load(rAdjGCN(_cfg(golden, r=0.5), golden_dataset(golden)), golden).eval()
    b = _load(LightGCN(_cfg(golden), golden_dataset(golden)), golden).eval()
    with torch.no_grad():
        assert_close(torch.cat(a.forward()), torch.cat(b.forward()), what="r=0.5")


def test_lightgcn_matches_pyg_form(golden, variants):
    """model/lgcn.py's LightGCN (PyG LGConv) on our kernels: same vectors as the torch.sparse form."""
    model = _load(LightGCN(_cfg(golden), golden_dataset(golden)), golden)
    _check(model, golden, variants, "pyg")


def test_rgcn_forward(golden, variants):
    ds = golden_dataset(golden)
    ds.favoriteUser, ds.favoriteItem = variants["fav_user"], variants["fav_item"]
    model = _load(RGCN(_cfg(golden), ds), golden).eval()
    assert model.graph.nnz == 2 * (len(golden["train_user"]) + len(variants["fav_user"]))
    with torch.no_grad():
        u, i = model.forward()
    assert_close(u, variants["rgcn_users"], what="rgcn users")
    assert_close(i, variants["rgcn_items"], what="rgcn items")
    # training step runs and still samples/evaluates on the purchase lists only
    model.train()
    l = model.stageOne(*batch(golden))
    assert np.isfinite(l.item())
    with pytest.raises(ValueError):
        RGCN(_cfg(golden), golden_dataset(golden))


def test_capped_sampler_bit_exact(golden, variants):
    ds = golden_dataset(golden)
    cap = int(variants["cap"])
    S = UniformSampleCapped(ds, limit=cap, seed=9, epoch=1, count=3000)
    assert S.dtype == torch.int64 and S.is_cuda
    assert np.array_equal(S.cpu().numpy(), variants["capped_philox_seed9_epoch1"])
    # default size is trainDataSize * TRAIN_ITERATIVE draws (ddp_lgcn.py:549); a huge cap is the plain sampler
    big = UniformSampleCapped(ds, limit=10 ** 9, seed=9, epoch=1)
    train = [ds.allPos[u] for u in range(ds.n_users)]
    full, _ = orc.uniform_sample_philox(train, ds.n_users, ds.m_items, 3 * ds.trainDataSize, seed=9, epoch=1)
    assert np.array_equal(big.cpu().numpy(), full)


def test_weighted_positive_sampler_bit_exact(golden, variants):
    from furusato_recommend_b200 import UniformSampling
    ds = golden_dataset(golden)
    flat = variants["weighted_probs_flat"]
    s = UniformSampling(ds, {"sample_pow": 0.5}, probs=flat)
    S = s.sample(seed=4, epoch=2, count=3000)
    # the device inverse-CDF table against numpy's per-user cumsum (np.random.choice's table)
    train = [ds.allPos[u] for u in range(ds.n_users)]
    ptr = np.concatenate([[0], np.cumsum([len(p) for p in train])])
    cdf_dev = s._cdf.cpu().numpy()
    cdfs = [cdf_dev[ptr[u]:ptr[u + 1]] for u in range(ds.n_users)]
    want = np.concatenate([orc.normalised_cdf(flat[ptr[u]:ptr[u + 1]]) for u in range(ds.n_users) if ptr[u + 1] > ptr[u]])
    assert np.abs(cdf_dev - want).max() < 1e-6
    # the kernel's decision procedure, bit-exact on its own table
    P, _ = orc.uniform_sample_philox(train, ds.n_users, ds.m_items, 3000, seed=4, epoch=2, pos_cdf=cdfs)
    assert np.array_equal(S.cpu().numpy(), P)
    if np.array_equal(cdf_dev, want):   # identical tables -> identical to the frozen vector
        assert np.array_equal(S.cpu().numpy(), variants["weighted_philox_seed4_epoch2"])
    # default probabilities: popularity^-pow, and sample_pow = 0 is the plain uniform sampler
    d = UniformSampling(ds, {"sample_pow": 0.5}).sample(seed=4, epoch=2, count=3000)
    pop = np.bincount(golden["train_item"], minlength=ds.m_items)
    u0 = UniformSampling(ds, {"sample_pow": 0}).sample(seed=4, epoch=2, count=3000)
    full, _ = orc.uniform_sample_philox(train, ds.n_users, ds.m_items, 3000, seed=4, epoch=2)
    assert np.array_equal(u0.cpu().numpy(), full)
    assert pop[d[:, 1].cpu().numpy()].mean() < pop[full[:, 1]].mean()


@pytest.mark.parametrize("storage", ["fp32", "bf16"])
def test_edge_dropout_matches_live_reference(golden, variants, storage):
    """Edge dropout (model/MF.py:158-192) through per-slot weights: forward with the mask the live
    reference drew, and the gradient through the TRANSPOSED dropped operator (reverse-entry weights)."""
    keep = float(variants["dropout_keep"])
    model = _load(LightGCN(_cfg(golden, dropout=1, keep_prob=keep, storage_dtype=storage), golden_dataset(golden)), golden)
    mask = variants["dropout_mask"]
    model.dropout_mask_fn = lambda n_entries: mask
    assert model.graph.entry_index()[1] == len(mask)
    rtol = 1e-5 if storage == "fp32" else 2e-2
    model.train()
    with torch.no_grad():
        u, i = model.computer()
    assert_close(u, variants["dropout_users"], rtol=rtol, what="dropout users")
    assert_close(i, variants["dropout_items"], rtol=rtol, what="dropout items")
    bu, bp, bn = batch(golden)
    loss, reg = model.bpr_loss(bu, bp, bn)
    assert abs(loss.item() - float(variants["dropout_loss"])) <= rtol * abs(float(variants["dropout_loss"]))
    model.optim.zero_grad()
    (loss + float(model.config["decay"]) * reg).backward()
    assert_close(model.all_embedding.weight.grad, variants["dropout_grad"], rtol=rtol, what="dropout grad")
    # eval mode never drops (MF.py:187-192), and the fused step draws a fresh mask per step
    model.eval()
    with torch.no_grad():
        eu, _ = model.computer()
    assert_close(eu, golden["computer_users"], rtol=rtol, what="eval users")
    model.train()
    model.dropout_mask_fn = None
    l1 = model.stageOne(bu, bp, bn).item()
    w1 = model._drop_fwd.clone()
    l2 = model.stageOne(bu, bp, bn).item()
    assert np.isfinite(l1) and np.isfinite(l2) and not torch.equal(w1, model._drop_fwd)
    frac = float((model._drop_fwd > 0).float().mean())
    assert abs(frac - keep) < 0.05


def test_ddp_shaped_model_with_a_caller_owned_optimizer(golden):
    """ddp_lgcn.py's call shape (forward(edge_index), OneEpoch(optimizer, ...), getUsersRating() -> tables) with
    a stock torch.optim.Adam the caller owns (ddp_lgcn.py:664): two steps reproduce the live reference's golden
    losses and table (autograd through torch.ops.lgcn_b200.propagate), and the model's own FusedAdam takes the
    fused path with the same result."""
    from furusato_recommend_b200 import DDPLightGCN
    cfg = _cfg(golden)
    bu, bp, bn = batch(golden)
    for own in (False, True):
        model = _load(DDPLightGCN(cfg, golden_dataset(golden)), golden).train()
        opt = model.optim if own else torch.optim.Adam(model.parameters(), lr=cfg["lr"])
        l1 = model.stageOne(opt, bu, bp, bn)
        l2 = model.stageOne(opt, bu, bp, bn)
        for l, key in ((l1, "step1_loss"), (l2, "step2_loss")):
            assert abs(float(l) - float(golden[key])) <= 1e-5 * abs(float(golden[key])), (own, key)
        assert float((model.all_embedding.weight.detach().cpu() - torch.from_numpy(golden["E2"])).abs().max()) < 2e-6
        model.eval()
        ux, ix = model.getUsersRating()                      # (user_x, item_x), ddp_lgcn.py:535-538
        fu, fi = model.forward(None)                         # edge_index accepted and ignored
        assert ux.shape == (int(golden["n_users"]), cfg["recdim"]) and torch.equal(ux, fu) and torch.equal(ix, fi)
        ep = model.train().OneEpoch(opt, bu.repeat(3)[:150], bp.repeat(3)[:150], bn.repeat(3)[:150])
        assert torch.isfinite(ep)
