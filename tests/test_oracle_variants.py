"""Oracle restatements of the f-4 variants (rAdjGCN, PyG-form LightGCN, RGCN, capped sampler)
against the vectors frozen from the live reference classes (oracle/make_golden_variants.py)."""
import numpy as np
import torch

from oracle import lgcn_oracle as orc


def _setup(golden):
    n, m = int(golden["n_users"]), int(golden["m_items"])
    d, K, B = (int(x) for x in golden["config"])
    lr, decay = (float(x) for x in golden["hyper"])
    edge = orc.directed_edges(n, golden["train_user"], golden["train_item"])
    batch = tuple(torch.from_numpy(golden[k]).long() for k in ("batch_users", "batch_pos", "batch_neg"))
    return n, m, K, lr, decay, edge, batch


def _check_model(golden, variants, tag, fwd):
    n, m, K, lr, decay, edge, (tu, tp, tn) = _setup(golden)
    w = torch.from_numpy(golden["E0"]).clone().requires_grad_(True)
    u, i = fwd(w, edge, K, n)
    assert np.array_equal(u.detach().numpy(), variants[f"{tag}_users"])
    assert np.array_equal(i.detach().numpy(), variants[f"{tag}_items"])
    loss, reg = orc.bpr_loss_from(u, i, w, n, tu, tp, tn)
    assert loss.item() == float(variants[f"{tag}_loss"]) and reg.item() == float(variants[f"{tag}_reg"])
    (loss + decay * reg).backward()
    assert np.allclose(w.grad.numpy(), variants[f"{tag}_grad"], rtol=0, atol=1e-9)
    # two Adam steps (stageOne of the live class)
    p = torch.nn.Parameter(torch.from_numpy(golden["E0"]).clone())
    opt = torch.optim.Adam([p], lr=lr)
    for step in (1, 2):
        opt.zero_grad()
        uu, ii = fwd(p, edge, K, n)
        l, r = orc.bpr_loss_from(uu, ii, p, n, tu, tp, tn)
        tot = l + decay * r
        tot.backward()
        opt.step()
        assert abs(tot.item() - float(variants[f"{tag}_step{step}_loss"])) < 1e-7
        assert np.allclose(p.detach().numpy(), variants[f"{tag}_E{step}"], rtol=0, atol=1e-7)


def test_radj_matches_live_reference(golden, variants):
    r = float(variants["r"])
    _check_model(golden, variants, "radj", lambda w, e, K, n: orc.radj_forward(w, e, K, n, r))


def test_radj_is_the_two_vector_scaling(golden, variants):
    """The identity the CUDA path relies on: rAdjConv == diag(deg^-(1-r)) A diag(deg^-r)."""
    n, m, K, *_ = _setup(golden)
    r = float(variants["r"])
    N = n + m
    edge = orc.directed_edges(n, golden["train_user"], golden["train_item"])
    deg = torch.bincount(edge[0], minlength=N).double()
    a = torch.where(deg > 0, deg.pow(-r), torch.zeros_like(deg))
    b = torch.where(deg > 0, deg.pow(-(1 - r)), torch.zeros_like(deg))
    A = torch.zeros(N, N, dtype=torch.float64).index_put_((edge[1], edge[0]), torch.ones(edge.shape[1], dtype=torch.float64),
                                                         accumulate=True)
    x = torch.from_numpy(golden["E0"]).double()
    want = orc.radj_conv(x.float(), edge, N, r).double()
    got = b[:, None] * (A @ (a[:, None] * x))
    assert (got - want).abs().max() / want.abs().max() < 1e-6


def test_pyg_form_matches_live_reference(golden, variants):
    _check_model(golden, variants, "pyg", orc.lgconv_forward)
    # and the PyG form agrees with the torch.sparse form pinned in tiny_ref.npz
    rel = np.abs(variants["pyg_users"] - golden["computer_users"]).max() / np.abs(golden["computer_users"]).max()
    assert rel < 1e-6


def test_rgcn_forward_frozen(golden, variants):
    n, m, K, *_ = _setup(golden)
    edge = orc.directed_edges(n, golden["train_user"], golden["train_item"], variants["fav_user"], variants["fav_item"])
    u, i = orc.lgconv_forward(torch.from_numpy(golden["E0"]), edge, K, n)
    assert np.array_equal(u.numpy(), variants["rgcn_users"]) and np.array_equal(i.numpy(), variants["rgcn_items"])


def test_capped_sampler(golden, variants, tiny_lists):
    train, _ = tiny_lists
    n, m = int(golden["n_users"]), int(golden["m_items"])
    cap = int(variants["cap"])
    np.random.seed(321)
    S = orc.capped_sample_mt(train, m, 3 * len(golden["train_user"]), cap)
    assert np.array_equal(S, variants["capped_mt_seed321"])
    assert np.bincount(S[:, 1]).max() == cap
    P = orc.capped_sample_philox(train, n, m, 3000, seed=9, epoch=1, limit=cap)
    assert np.array_equal(P, variants["capped_philox_seed9_epoch1"])
    assert np.bincount(P[:, 1]).max() == cap
    # the cap only removes rows of the uncapped Philox sample, in order
    full, _ = orc.uniform_sample_philox(train, n, m, 3000, seed=9, epoch=1)
    it = iter(full.tolist())
    assert all(any(row == cand for cand in it) for row in P.tolist())


def test_weighted_positive_sampler(golden, variants, tiny_lists):
    """UniformSampling with sample_pow != 0 (negative_sample.py:53-56): decision procedure under the
    reference's RNG, and the Philox restatement on fp32 inverse-CDF tables."""
    train, _ = tiny_lists
    n, m = int(golden["n_users"]), int(golden["m_items"])
    flat = variants["weighted_probs_flat"]
    ptr = np.concatenate([[0], np.cumsum([len(p) for p in train])])
    probs = [flat[ptr[u]:ptr[u + 1]] for u in range(n)]
    np.random.seed(77)
    W = orc.weighted_sample_mt(train, probs, n, m, 4000)
    assert np.array_equal(W, variants["weighted_mt_seed77"])
    cdfs = [orc.normalised_cdf(p) if len(p) else np.zeros(0, np.float32) for p in probs]
    P, _ = orc.uniform_sample_philox(train, n, m, 3000, seed=4, epoch=2, pos_cdf=cdfs)
    assert np.array_equal(P, variants["weighted_philox_seed4_epoch2"])
    # same users and negatives' acceptance rule as the uniform sampler; only the positive pick moves
    U, _ = orc.uniform_sample_philox(train, n, m, 3000, seed=4, epoch=2)
    assert np.array_equal(P[:, 0], U[:, 0]) and not np.array_equal(P[:, 1], U[:, 1])
    # rare items are preferred: mean popularity of the weighted positives is lower
    pop = np.bincount(golden["train_item"], minlength=m)
    assert pop[P[:, 1]].mean() < pop[U[:, 1]].mean()


def test_edge_dropout_matches_live_reference(golden, variants):
    """model/MF.py:158-192 with the mask the live reference drew under torch.manual_seed(99)."""
    n, m, K, lr, decay, edge, (tu, tp, tn) = _setup(golden)
    keep = float(variants["dropout_keep"])
    g = orc.sparse_graph(n, m, golden["train_user"], golden["train_item"])
    torch.manual_seed(99)
    mask = (torch.rand(len(g.values())) + keep).int().bool()       # MF.py:162-163
    assert np.array_equal(mask.numpy(), variants["dropout_mask"])
    gd = orc.dropout_graph(g, mask, keep)
    w = torch.from_numpy(golden["E0"]).clone().requires_grad_(True)
    u, i = orc.computer(w, gd, K, n)
    assert np.array_equal(u.detach().numpy(), variants["dropout_users"])
    loss, reg = orc.bpr_loss(w, gd, K, n, tu, tp, tn)
    assert loss.item() == float(variants["dropout_loss"]) and reg.item() == float(variants["dropout_reg"])
    (loss + decay * reg).backward()
    assert np.allclose(w.grad.numpy(), variants["dropout_grad"], rtol=0, atol=1e-9)
