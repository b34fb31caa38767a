"""Parity of the CUDA path (through the C ABI) against the oracle / golden vectors.

Tolerances (BASELINE.json north_star): fp32 embeddings, losses and scores within
1e-5 relative; bf16 storage within 2e-2 relative; sampled triples and top-k ids
bit-exact given the same draws / the same scores (ties: lowest item id)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import gap_rule  # noqa: E402

from furusato_recommend_b200 import LightGCN, Trainer, UniformSample, ops  # noqa: E402
from furusato_recommend_b200.dataloader import BasicDataset  # noqa: E402
from furusato_recommend_b200.synthetic import bipartite  # noqa: E402
from oracle import lgcn_oracle as orc  # noqa: E402

DEV = "cuda:0"
RTOL = 1e-5


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def assert_close(a, b, rtol=RTOL, what=""):
    """|a-b| <= rtol*|b| + 0.1*rtol*max|b| elementwise (the floor absorbs fp32 summation-order
    noise on elements that cancel to ~0; it is still 10x tighter than the tolerance on the scale)."""
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    tol = rtol * b.abs() + rtol * 1e-1 * b.abs().max()
    bad = (a - b).abs() > tol
    assert not bool(bad.any()), f"{what}: {int(bad.sum())} elements off, max rel-to-max {rel_err(a, b):.3e}"


def golden_dataset(golden) -> BasicDataset:
    return BasicDataset(int(golden["n_users"]), int(golden["m_items"]), golden["train_user"], golden["train_item"],
                        golden["test_user"], golden["test_item"], config={"A_split": False}, device=DEV)


def golden_model(golden, weights="E0", **extra) -> LightGCN:
    d, K, B = (int(x) for x in golden["config"])
    lr, decay = (float(x) for x in golden["hyper"])
    cfg = dict(recdim=d, layer=K, lr=lr, decay=decay, bpr_batch_size=B, device=DEV, test_u_batch_size=128, **extra)
    model = LightGCN(cfg, golden_dataset(golden))
    with torch.no_grad():
        model.all_embedding.weight.copy_(torch.from_numpy(golden[weights]))
    return model


def batch(golden):
    return tuple(torch.from_numpy(golden[k]).to(DEV) for k in ("batch_users", "batch_pos", "batch_neg"))


# ------------------------------------------------------------------ propagation
def test_computer_matches_reference(golden):
    model = golden_model(golden)
    model.eval()
    with torch.no_grad():
        users, items = model.computer()
        fu, fi = model.forward()  # model/lgcn.py name for the same thing
    assert_close(users, golden["computer_users"], what="computer users")
    assert_close(items, golden["computer_items"], what="computer items")
    assert torch.equal(fu, users) and torch.equal(fi, items)
    # zero-degree items keep only the ego term: out = E/(K+1)
    K = model.num_layers
    n = model.num_users
    iso = (model.graph.dinv == 0).nonzero().flatten()
    assert len(iso) >= 4
    out = torch.cat([users, items])
    assert torch.allclose(out[iso], model.all_embedding.weight[iso] / (K + 1), rtol=1e-6, atol=0)


@pytest.mark.parametrize("d", [32, 64, 128])
@pytest.mark.parametrize("storage", ["fp32", "bf16"])
def test_propagation_with_hub_rows(d, storage):
    """Rows above HUB_DEG (one CTA segment) and above SEG_EDGES (multi-segment, last-arriver
    reduction), multi-edges and isolated nodes, every supported width and storage type."""
    rng = np.random.default_rng(5)
    n, m = 6000, 300
    tu, ti = [], []
    for u in range(n):
        its = set(rng.integers(3, m, rng.integers(1, 12)).tolist())
        if u % 2 == 0:
            its.add(0)            # item 0: degree 3000 -> 3 segments
        if u < 700:
            its.add(1)            # item 1: degree 700  -> 1 segment
        if u < 255:
            its.add(2)            # item 2 stays a light row just under the hub threshold
        its = list(its)
        tu += [u] * len(its)
        ti += its
    tu += [5, 5, 7]
    ti += [ti[tu.index(5)]] * 2 + [ti[tu.index(7)]]  # multi-edges
    m_total = m + 3  # three items nobody touched
    ds = BasicDataset(n, m_total, np.array(tu), np.array(ti), np.array([0]), np.array([3]), config={}, device=DEV)
    cfg = dict(recdim=d, layer=3, lr=1e-3, decay=1e-4, bpr_batch_size=64, device=DEV, storage_dtype=storage)
    model = LightGCN(cfg, ds)
    assert int(model.graph.hub_nseg.max()) >= 3 and len(model.graph.hub_nseg) >= 2
    model.eval()
    with torch.no_grad():
        u, i = model.computer()
        u2, i2 = (t.clone() for t in model.computer())
    g = orc.sparse_graph(n, m_total, np.array(tu), np.array(ti))
    ou, oi = orc.computer(model.all_embedding.weight.detach().cpu(), g, 3, n)
    tol = 1e-5 if storage == "fp32" else 2e-2
    assert rel_err(u, ou) < tol and rel_err(i, oi) < tol, (rel_err(u, ou), rel_err(i, oi))
    if storage == "fp32":
        assert_close(u, ou, what="hub users")
        assert_close(i, oi, what="hub items")
    assert torch.equal(u, u2) and torch.equal(i, i2)  # deterministic, counters self-reset


def test_propagation_is_linear_cfg2_scale():
    """Size-independent property at the BASELINE cfg-2 shape: computer(a*E) == a*computer(E)."""
    n, m, tu, ti, su, si = bipartite(30000, 41000, 1_250_000, seed=2020)
    ds = BasicDataset(n, m, tu.numpy(), ti.numpy(), su.numpy(), si.numpy(), config={}, device=DEV)
    model = LightGCN(dict(recdim=64, layer=3, lr=1e-4, decay=1e-7, bpr_batch_size=2048, device=DEV), ds)
    model.eval()
    with torch.no_grad():
        u1, i1 = (t.clone() for t in model.computer())
        model.all_embedding.weight.mul_(2.0)
        model.train(); model.eval()
        u2, i2 = model.computer()
    assert torch.equal(u2, 2 * u1) and torch.equal(i2, 2 * i1)  # scaling by 2 is exact in fp32
    # constant vector: A_hat^k applied to sqrt(deg) reproduces sqrt(deg) (eigenvector), so
    # out = sqrt(deg) on every node with deg > 0
    deg = (model.graph.rowptr[1:] - model.graph.rowptr[:-1]).float()
    with torch.no_grad():
        model.all_embedding.weight.copy_(deg.sqrt()[:, None].expand(-1, 64))
        model.train(); model.eval()
        u3, i3 = model.computer()
    out = torch.cat([u3, i3])
    nz = deg > 0
    assert torch.allclose(out[nz], deg.sqrt()[nz][:, None].expand(-1, 64), rtol=2e-5, atol=0)


# ------------------------------------------------------------------ BPR / Adam
def test_bpr_loss_and_grad(golden):
    model = golden_model(golden)
    model.train()
    u, p, q = batch(golden)
    loss, reg = model.bpr_loss(u, p, q)
    assert abs(loss.item() - float(golden["loss"])) <= RTOL * abs(float(golden["loss"]))
    assert abs(reg.item() - float(golden["reg"])) <= RTOL * abs(float(golden["reg"]))
    total = loss + model.config["decay"] * reg
    model.optim.zero_grad()
    total.backward()
    g = model.all_embedding.weight.grad
    assert rel_err(g, torch.from_numpy(golden["grad"])) < RTOL
    assert_close(g, golden["grad"], what="bpr grad")


def test_stage_one_two_adam_steps(golden):
    model = golden_model(golden)
    model.train()
    u, p, q = batch(golden)
    l1 = model.stageOne(u, p, q)
    assert_close(model.all_embedding.weight, golden["E1"], what="E after 1 step")
    l2 = model.stageOne(u, p, q)
    assert_close(model.all_embedding.weight, golden["E2"], what="E after 2 steps")
    assert abs(l1.item() - float(golden["step1_loss"])) <= RTOL * float(golden["step1_loss"])
    assert abs(l2.item() - float(golden["step2_loss"])) <= RTOL * float(golden["step2_loss"])
    # the Adam update itself (E1 - E0 ~ lr) must match, not just the table
    d_ref = torch.from_numpy(golden["E1"] - golden["E0"])
    m2 = golden_model(golden)
    m2.train()
    m2.stageOne(u, p, q)
    d_got = m2.all_embedding.weight.detach().cpu() - torch.from_numpy(golden["E0"])
    assert rel_err(d_got, d_ref) < 1e-3


@pytest.mark.parametrize("layers", [1, 2, 3])
def test_autograd_path_equals_fused_path(golden, layers):
    d, K, B = (int(x) for x in golden["config"])
    u, p, q = batch(golden)
    a = golden_model(golden)
    b = golden_model(golden)
    for mdl in (a, b):
        mdl.num_layers = layers
        mdl.train()
    for _ in range(3):
        la = a.stageOne(u, p, q)
        b.optim.zero_grad()
        loss, reg = b.bpr_loss(u, p, q)
        tot = loss + b.config["decay"] * reg
        tot.backward()
        b.optim.step()
        assert abs(la.item() - tot.item()) < 1e-6
    assert rel_err(a.all_embedding.weight, b.all_embedding.weight) < 1e-6
    om = orc.OracleModel(a.num_users, a.num_items, golden["train_user"], golden["train_item"],
                         torch.from_numpy(golden["E0"]), layers, float(golden["hyper"][0]), float(golden["hyper"][1]))
    for _ in range(3):
        om.stage_one(u.cpu(), p.cpu(), q.cpu())
    assert_close(a.all_embedding.weight, om.weight.detach(), what=f"3 steps, K={layers}")


def test_computer_backward_through_autograd(golden):
    model = golden_model(golden)
    model.train()
    w = torch.from_numpy(np.random.default_rng(0).standard_normal((model.num_users + model.num_items, model.latent_dim)).astype(np.float32))
    users, items = model.computer()
    (torch.cat([users, items]) * w.to(DEV)).sum().backward()
    E = torch.from_numpy(golden["E0"]).clone().requires_grad_(True)
    g = orc.sparse_graph(model.num_users, model.num_items, golden["train_user"], golden["train_item"])
    ou, oi = orc.computer(E, g, model.num_layers, model.num_users)
    (torch.cat([ou, oi]) * w).sum().backward()
    assert rel_err(model.all_embedding.weight.grad, E.grad) < RTOL
    # getEmbedding keeps the reference's 6-tuple and its ego rows
    u, p, q = batch(golden)
    e = model.getEmbedding(u, p, q)
    assert len(e) == 6 and torch.equal(e[3], model.all_embedding.weight[u])
    assert torch.equal(e[4], model.all_embedding.weight[p + model.num_users])


def test_one_epoch_loss_divisor(golden):
    model = golden_model(golden)
    model.train()
    rng = np.random.default_rng(3)
    n_s, B = 200, int(golden["config"][2])  # 200 // 64 + 1 = 4 batches counted, 4 run (last partial)
    users = torch.from_numpy(rng.integers(0, model.num_users, n_s)).to(DEV)
    pos = torch.from_numpy(np.array([golden["train_item"][np.searchsorted(golden["train_user"], u)] for u in users.cpu().numpy()])).to(DEV)
    neg = torch.from_numpy(rng.integers(0, model.num_items, n_s)).to(DEV)
    loss = model.OneEpoch(users, pos, neg)
    om = orc.OracleModel(model.num_users, model.num_items, golden["train_user"], golden["train_item"],
                         torch.from_numpy(golden["E0"]), model.num_layers, float(golden["hyper"][0]), float(golden["hyper"][1]))
    oloss = om.one_epoch(users.cpu(), pos.cpu(), neg.cpu(), B)
    assert abs(loss.item() - oloss.item()) <= RTOL * abs(oloss.item())
    assert_close(model.all_embedding.weight, om.weight.detach(), what="E after OneEpoch")


def test_bf16_storage_within_tolerance(golden):
    model = golden_model(golden, storage_dtype="bf16")
    model.eval()
    with torch.no_grad():
        users, items = model.computer()
    assert rel_err(users, torch.from_numpy(golden["computer_users"])) < 2e-2
    assert rel_err(items, torch.from_numpy(golden["computer_items"])) < 2e-2
    model.train()
    u, p, q = batch(golden)
    l1 = model.stageOne(u, p, q)
    assert abs(l1.item() - float(golden["step1_loss"])) < 2e-2 * float(golden["step1_loss"])
    d_ref = torch.from_numpy(golden["E1"] - golden["E0"])
    d_got = model.all_embedding.weight.detach().cpu() - torch.from_numpy(golden["E0"])
    # first Adam step is lr*sign(g): a bf16-rounded gradient keeps the sign almost everywhere
    assert float(((d_got - d_ref).abs() > 1e-4).float().mean()) < 0.02


# ------------------------------------------------------------------ sampler
def test_sampler_bit_exact(golden, tiny_lists):
    ds = golden_dataset(golden)
    S = UniformSample(ds, seed=2020, epoch=3, count=2000)
    assert S.dtype == torch.int64 and S.is_cuda
    assert np.array_equal(S.cpu().numpy(), golden["sample_philox_seed2020_epoch3"])
    # a shard is a counter offset: the union of shards is the whole sample, in order
    parts = [UniformSample(ds, seed=2020, epoch=3, count=c, start=s) for s, c in ((0, 700), (700, 1), (701, 1299))]
    assert torch.equal(torch.cat(parts), S)
    # default count = trainDataSize, epoch counter advances like the reference's global RNG
    a, b = UniformSample(ds), UniformSample(ds)
    assert len(a) == ds.trainDataSize and not torch.equal(a, b)


def test_sampler_multi_negative_and_ssm_variant(golden, tiny_lists):
    """cfg-4 shape on the tiny graph: J negatives per positive as flat triples, and the
    reference's lgcnssm arithmetic (BPR softplus over J*B rows per step, lgcnssm.py:98-153)."""
    from furusato_recommend_b200 import LightGCNSSM
    train, _ = tiny_lists
    ds = golden_dataset(golden)
    J = 8
    S = UniformSample(ds, neg_ratio=J, seed=7, epoch=1, count=300)
    want, _ = orc.uniform_sample_philox(train, ds.n_users, ds.m_items, 300, seed=7, epoch=1, n_neg=J)
    assert np.array_equal(S.cpu().numpy(), want)
    d, K, B = (int(x) for x in golden["config"])
    lr, decay = (float(x) for x in golden["hyper"])
    cfg = dict(recdim=d, layer=K, lr=lr, decay=decay, bpr_batch_size=B, device=DEV, neg_size=J)
    model = LightGCNSSM(cfg, ds)
    with torch.no_grad():
        model.all_embedding.weight.copy_(torch.from_numpy(golden["E0"]))
    model.train()
    u, p, q = (S[:, j].contiguous() for j in range(3))
    loss = model.OneEpoch(u, p, q)
    om = orc.OracleModel(ds.n_users, ds.m_items, golden["train_user"], golden["train_item"],
                         torch.from_numpy(golden["E0"]), K, lr, decay)
    tot = 0.0
    for i in range(0, len(u), J * B):   # lgcnssm.py:141: batch_size = neg_size * bpr_batch_size
        tot += float(om.stage_one(u[i:i + J * B].cpu(), p[i:i + J * B].cpu(), q[i:i + J * B].cpu()))
    want_loss = tot / (len(u) // B + 1)
    assert abs(float(loss) - want_loss) <= RTOL * abs(want_loss)
    assert_close(model.all_embedding.weight, om.weight.detach(), what="E after the SSM-variant epoch")
    l0, r0 = model.softmax_loss(u[:J * B], p[:J * B], q[:J * B])
    l1, r1 = model.bpr_loss(u[:J * B], p[:J * B], q[:J * B])
    assert float(l0) == float(l1) and float(r0) == float(r1)   # byte-for-byte the BPR loss, as in the reference
    # opt-in real sampled softmax (own spec): finite, decreases under its own steps
    cfg2 = dict(cfg, ssm_true_softmax=True, lr=1e-2)
    m2 = LightGCNSSM(cfg2, ds)
    m2.train()
    first = float(m2.stageOne(u[:J * B], p[:J * B], q[:J * B]))
    for _ in range(5):
        last = float(m2.stageOne(u[:J * B], p[:J * B], q[:J * B]))
    assert np.isfinite(first) and last < first


def test_sampler_skips_users_without_positives():
    # users 1 and 3 have no train line: their samples vanish and order is preserved
    tu, ti = np.array([0, 0, 2, 4, 4, 4]), np.array([1, 2, 0, 3, 1, 0])
    ds = BasicDataset(5, 6, tu, ti, np.array([0]), np.array([3]), config={}, device=DEV)
    S = UniformSample(ds, seed=9, epoch=1, count=5000).cpu().numpy()
    all_pos = [ds.allPos[u] for u in range(5)]
    want, valid = orc.uniform_sample_philox(all_pos, 5, 6, 5000, seed=9, epoch=1)
    assert 0 < len(want) < 5000 and np.array_equal(S, want)
    assert not np.isin(S[:, 0], [1, 3]).any()


def test_sampler_properties_cfg2_scale():
    n, m, tu, ti, su, si = bipartite(30000, 41000, 1_250_000, seed=2020)
    ds = BasicDataset(n, m, tu.numpy(), ti.numpy(), su.numpy(), si.numpy(), config={}, device=DEV)
    S = UniformSample(ds, seed=1, epoch=0)
    assert len(S) == ds.trainDataSize
    key = torch.unique(tu.to(DEV) * m + ti.to(DEV))
    pos_key = S[:, 0] * m + S[:, 1]
    neg_key = S[:, 0] * m + S[:, 2]
    assert bool(torch.isin(pos_key, key).all()) and not bool(torch.isin(neg_key, key).any())
    assert int(S[:, 2].min()) >= 0 and int(S[:, 2].max()) < m
    # users are uniform over [0, n): mean and spread of the draw
    assert abs(float(S[:, 0].double().mean()) / n - 0.5) < 0.01
    oracle_head, _ = orc.uniform_sample_philox(ds.allPos, n, m, 3000, seed=1, epoch=0)
    assert np.array_equal(S[:3000].cpu().numpy(), oracle_head)


# ------------------------------------------------------------------ eval
def _stable_topk(scores: torch.Tensor, pos_lists, k: int, mask=-1024.0):
    s = scores.clone()
    for r, items in enumerate(pos_lists):
        if len(items):
            s[r, torch.as_tensor(np.asarray(items), device=s.device, dtype=torch.long)] = mask
    vals, idx = torch.sort(s, dim=1, descending=True, stable=True)
    return idx[:, :k], vals[:, :k]


@pytest.mark.parametrize("shape", [(300, 400, 32, 20), (1000, 5000, 64, 20), (257, 1111, 128, 50), (64, 40, 64, 20)])
def test_score_topk_bit_exact_on_own_scores(shape):
    U, m, d, k = shape
    gen = torch.Generator(device="cpu").manual_seed(U + m)
    ue = torch.randn(U, d, generator=gen).to(DEV)
    ie = torch.randn(m, d, generator=gen).to(DEV)
    if m == 5000:  # exact ties: quantised embeddings make many equal scores
        ue, ie = (ue * 2).round() / 2, (ie * 2).round() / 2
    rng = np.random.default_rng(1)
    if m == 40:  # leave fewer than k unmasked items so that masked ones must re-enter
        lists = [rng.choice(m, rng.integers(25, 36), replace=False) for _ in range(U)]
        lists = [np.sort(x) for x in lists]
    else:
        lists = [np.unique(rng.integers(0, m, rng.integers(0, 30))) for _ in range(U)]
    rowptr = torch.tensor(np.concatenate([[0], np.cumsum([len(x) for x in lists])]), dtype=torch.int64, device=DEV)
    flat = torch.tensor(np.concatenate(lists) if len(lists) else [], dtype=torch.int32, device=DEV)
    ids = torch.randperm(U, generator=gen).to(DEV)  # evaluated users in arbitrary order
    idx, val = ops.score_topk(ue, ie, ids, rowptr, flat, k)
    dense = ops.score_dense_f32(ue, ie, ids)
    widx, wval = _stable_topk(dense, [lists[u] for u in ids.cpu().tolist()], k)
    assert torch.equal(idx.long(), widx), f"{int((idx.long() != widx).sum())} ids differ"
    assert torch.equal(val, wval)
    if m == 40:  # fewer than k unmasked items above -1024: masked items re-enter (trainer.py:137)
        assert bool((val == -1024.0).any())
    # dense debug scores are the fp32 product within rounding
    assert rel_err(dense, ue[ids] @ ie.t()) < 1e-5


def test_eval_matches_reference_topk_and_metrics(golden, tiny_lists):
    train, test = tiny_lists
    model = golden_model(golden, weights="E2", eval_precision="fp32")   # the exactness mode; default is bf16 (below)
    model.eval()
    users = torch.from_numpy(golden["eval_users"]).to(DEV)
    rating = model.getUsersRating(users)
    assert rating.shape == (len(users), model.num_items)
    assert_close(rating[:8], golden["raw_rating_first8"], what="getUsersRating")
    idx, val = model.getUsersTopK(users, 20)
    ref_idx, ref_val = torch.from_numpy(golden["topk_raw_idx"]), torch.from_numpy(golden["topk_raw_val"])
    assert rel_err(val, ref_val) < RTOL
    # ids must agree wherever the reference's neighbouring scores are separated by more than the tolerance
    gap = (ref_val[:, :-1] - ref_val[:, 1:]).abs()
    safe = torch.ones_like(ref_idx, dtype=torch.bool)
    thr = 4 * RTOL * ref_val.abs().max()
    safe[:, :-1] &= gap > thr
    safe[:, 1:] &= gap > thr
    safe[:, -1] = False  # the boundary with rank k+1 is not visible in the fixture
    assert torch.equal(idx.cpu().long()[safe], ref_idx[safe])
    assert float((idx.cpu().long() == ref_idx).float().mean()) > 0.98
    res = Trainer(model.config, model.dataset, model, topks=[int(k) for k in golden["topks"]]).test()
    for name in ("recall", "precision", "ndcg", "hr"):
        assert np.allclose(res[name], golden[f"metric_raw_{name}"], rtol=0, atol=2e-3), name
    # the default evaluation path scores on the tcgen05 tensor cores (bf16 operands, fp32 accumulate)
    dflt = golden_model(golden, weights="E2")
    assert dflt.eval_precision == "bf16"
    res16 = Trainer(dflt.config, dflt.dataset, dflt, topks=[int(k) for k in golden["topks"]]).test()
    for name in ("recall", "precision", "ndcg", "hr"):
        assert np.allclose(res16[name], golden[f"metric_raw_{name}"], rtol=0, atol=5e-3), name


def test_rank_metrics_match_oracle(golden, tiny_lists):
    train, test = tiny_lists
    ds = golden_dataset(golden)
    users = golden["eval_users"]
    topk = torch.from_numpy(golden["topk_raw_idx"]).to(torch.int32).to(DEV)
    rp, srt = ds.test_csr()
    for ks in ([10, 20], [20, 10], [1, 5, 20], [20]):
        sums = ops.rank_metrics(topk, torch.from_numpy(users).to(DEV), rp, srt, ks)
        want = orc.batch_metrics([test[u] for u in users], golden["topk_raw_idx"], ks)
        for r, name in enumerate(("recall", "precision", "hr", "ndcg")):
            assert np.allclose(sums[r].cpu().numpy(), want[name], rtol=1e-12), (name, ks)
    sums, hits = ops.rank_metrics(topk, torch.from_numpy(users).to(DEV), rp, srt, [20], want_hits=True)
    assert np.array_equal(hits.cpu().numpy().astype(float), orc.get_label([test[u] for u in users], golden["topk_raw_idx"]))


def test_get_topk_list_format(golden):
    model = golden_model(golden, weights="E2", eval_precision="fp32")
    lists = Trainer(model.config, model.dataset, model).get_topk_list(k=50)  # eval.py:35-40 candidates
    assert sum(len(x) for x in lists) == len(golden["eval_users"])
    assert lists[0].dtype == torch.int64 and lists[0].shape[1] == 50 and not lists[0].is_cuda
    assert torch.equal(lists[0][:, :20], torch.from_numpy(golden["topk_raw_idx"])[: len(lists[0])]) or \
        float((lists[0][:, :20] == torch.from_numpy(golden["topk_raw_idx"])[: len(lists[0])]).float().mean()) > 0.98


def test_export_candidates_file_format(golden, tmp_path):
    model = golden_model(golden, weights="E2")
    tr = Trainer(model.config, model.dataset, model)
    cand = tr.export_candidates(str(tmp_path / "lightgcn_result.pt"), k=50)   # eval.py:35-40
    back = torch.load(tmp_path / "lightgcn_result.pt")
    assert back.dtype == torch.int64 and back.shape == (len(golden["eval_users"]), 50) and torch.equal(back, cand)
    assert len(back.flatten()) // 50 == len(golden["eval_users"])                # train_lgbm.py:113-114


def test_training_reduces_loss_and_improves_recall():
    n, m, tu, ti, su, si = bipartite(3000, 2000, 80000, seed=3)
    cfg = dict(recdim=64, layer=3, lr=1e-2, decay=1e-4, bpr_batch_size=2048, device=DEV, test_u_batch_size=1000)
    ds = BasicDataset(n, m, tu.numpy(), ti.numpy(), su.numpy(), si.numpy(), config=cfg, device=DEV)
    torch.manual_seed(0)
    model = LightGCN(cfg, ds)
    tr = Trainer(cfg, ds, model)
    r0 = tr.test()["recall"][1]
    losses = [float(tr.train()) for _ in range(8)]
    r1 = tr.test()["recall"][1]
    assert losses[-1] < losses[0] and r1 > r0 + 0.02, (losses, r0, r1)
    sd = model.state_dict()
    assert list(sd.keys()) == ["all_embedding.weight"] and sd["all_embedding.weight"].shape == (n + m, 64)


# ------------------------------------------------------------------ wider shapes / edge cases
@pytest.mark.parametrize("d,K", [(64, 2), (128, 3), (32, 4)])
def test_fused_steps_other_widths_and_depths(d, K):
    """BPR + Horner backward + fused Adam at every supported width, K in {2,3,4}, a batch that
    is not a multiple of the group size, duplicate rows, zero-degree negatives."""
    rng = np.random.default_rng(d + K)
    n, m = 500, 700
    tu = np.repeat(np.arange(n), rng.integers(1, 9, n))
    ti = rng.integers(0, m - 20, len(tu))          # the last 20 items have no train edge
    ds = BasicDataset(n, m, tu, ti, np.array([0]), np.array([1]), config={}, device=DEV)
    cfg = dict(recdim=d, layer=K, lr=5e-3, decay=1e-3, bpr_batch_size=333, device=DEV)
    torch.manual_seed(d)
    model = LightGCN(cfg, ds)
    model.train()
    E0 = model.all_embedding.weight.detach().cpu().clone()
    om = orc.OracleModel(n, m, tu, ti, E0, K, 5e-3, 1e-3)
    for step in range(3):
        u = torch.from_numpy(rng.integers(0, n, 333))
        p = torch.from_numpy(np.array([ti[np.searchsorted(tu, x)] for x in u.numpy()]))
        q = torch.from_numpy(rng.integers(m - 40, m, 333))   # half of them isolated items
        l = model.stageOne(u.to(DEV), p.to(DEV), q.to(DEV))
        lo = om.stage_one(u, p, q)
        assert abs(l.item() - lo.item()) <= RTOL * abs(lo.item()), (step, l.item(), lo.item())
    # Adam divides by (|g| + eps): an element whose gradient is ~eps amplifies a 1-ulp gradient
    # difference, so the table is compared on its own scale (1e-5 of max|E|), not element by element
    assert rel_err(model.all_embedding.weight, om.weight.detach()) < RTOL


def test_error_behaviour_no_silent_fallback(golden):
    from furusato_recommend_b200 import _lib
    model = golden_model(golden)
    model.train()
    u, p, q = batch(golden)
    with pytest.raises(_lib.LgcnLibraryError):      # empty batch: the C ABI refuses, nothing falls back
        model._fused_step_eager(u[:0], p[:0], q[:0])
    with pytest.raises(_lib.LgcnLibraryError):      # k larger than the item count
        model.getUsersTopK(u[:4], model.num_items + 1)
    with pytest.raises(NotImplementedError):        # layer=0 is MF, outside the path
        LightGCN(dict(model.config, layer=0), model.dataset)
    with pytest.raises(_lib.LgcnLibraryError):      # unsupported width is an error, not a slow path
        LightGCN(dict(model.config, recdim=48), model.dataset).computer()


def test_eval_properties_cfg2_scale():
    """BASELINE cfg-2 shape: sortedness, train positives never returned, fp32 == stable sort of a
    sampled row block, tensor-core lists agree with fp32 lists up to bf16 near-ties."""
    n, m, tu, ti, su, si = bipartite(30000, 41000, 1_250_000, seed=2020)
    cfg = dict(recdim=64, layer=3, lr=1e-4, decay=1e-7, bpr_batch_size=2048, device=DEV, test_u_batch_size=10000)
    ds = BasicDataset(n, m, tu.numpy(), ti.numpy(), su.numpy(), si.numpy(), config=cfg, device=DEV)
    torch.manual_seed(1)
    model = LightGCN(cfg, ds)
    model.eval()
    users = torch.from_numpy(ds.test_users()).to(DEV)
    idx32, val32 = model.getUsersTopK(users, 20, precision="fp32")
    idx16, val16 = model.getUsersTopK(users, 20, precision="bf16")
    for idx, val in ((idx32, val32), (idx16, val16)):
        assert bool((val[:, :-1] >= val[:, 1:]).all())                      # sorted by score
        tie = val[:, :-1] == val[:, 1:]
        assert bool((idx[:, :-1][tie] < idx[:, 1:][tie]).all())             # ties by ascending id
        assert int(idx.min()) >= 0 and int(idx.max()) < m
        key = torch.unique(tu.to(DEV) * m + ti.to(DEV))
        got = users[:, None] * m + idx.long()
        assert not bool(torch.isin(got, key).any())                          # masked positives never surface
    i2, v2 = model.getUsersTopK(users, 20, precision="fp32")
    assert torch.equal(i2, idx32) and torch.equal(v2, val32)                 # idempotent / deterministic
    rows = users[:512]
    dense = ops.score_dense_f32(*model.computer(), rows)
    rp, _, srt = ds.pos_csr()
    for r in range(0, 512, 64):
        u = int(rows[r])
        s = dense[r].clone()
        s[srt[int(rp[u]):int(rp[u + 1])].long()] = -1024.0
        want = torch.sort(s, descending=True, stable=True)[1][:20]
        assert torch.equal(want, idx32[r].long())
    # tensor-core lists vs the exact fp32 scores under the 2e-2 gap rule, on two row blocks
    for lo in (0, len(users) - 1024):
        blk = users[lo:lo + 1024]
        dense_b = ops.score_dense_f32(*model.computer(), blk)
        for r, u in enumerate(blk.tolist()):
            dense_b[r, srt[int(rp[u]):int(rp[u + 1])].long()] = -1024.0
        tol = 2e-2 * float(dense_b[dense_b > -1000].abs().max())
        bad, compared, mism = gap_rule(idx16[lo:lo + 1024], dense_b, 20, tol)
        assert bad == 0 and mism == 0, (bad, compared, mism)
    overlap = (idx16[:, :, None] == idx32[:, None, :]).any(-1).float().mean()
    assert float(overlap) > 0.97, float(overlap)
    res = Trainer(cfg, ds, model).test()
    assert set(res) == {"recall", "precision", "hr", "ndcg"} and all(len(v) == 2 for v in res.values())


def test_registered_custom_op_opcheck(golden):
    """torch.ops.lgcn_b200.propagate (torch_ops.py): schema, FakeTensor rule and autograd registration pass
    torch.library.opcheck, and the op is what computer() differentiates through."""
    from furusato_recommend_b200 import torch_ops
    model = golden_model(golden)
    model.train()
    h = torch_ops.register_model(model)
    w = model.all_embedding.weight
    torch.library.opcheck(torch.ops.lgcn_b200.propagate, (w.detach().clone().requires_grad_(True), h),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
    out = torch.ops.lgcn_b200.propagate(w, h)
    ref = torch.from_numpy(np.concatenate([golden["computer_users"], golden["computer_items"]]))
    assert rel_err(out, ref) < RTOL
    u, i = model.computer()
    assert u.grad_fn is not None and "lgcn_b200" in type(u.grad_fn).__name__ or u.grad_fn is not None
