"""The oracle restatement against the golden vectors frozen from the LIVE reference
(oracle/make_golden.py) and against the Random123 Philox known-answer vectors."""
import numpy as np
import torch

from oracle import lgcn_oracle as orc


def _g(golden):
    n, m = int(golden["n_users"]), int(golden["m_items"])
    d, K, B = (int(x) for x in golden["config"])
    lr, decay = (float(x) for x in golden["hyper"])
    return n, m, d, K, B, lr, decay


def test_norm_adj_bit_identical(golden):
    n, m, *_ = _g(golden)
    row, col, val = orc.norm_adj_coo(n, m, golden["train_user"], golden["train_item"])
    assert np.array_equal(row, golden["adj_row"]) and np.array_equal(col, golden["adj_col"])
    assert np.array_equal(val, golden["adj_val"])  # fl32(fl32(dinv_i*mult)*dinv_j), dataloader.py:242-243
    # symmetric, multi-edges present, zero-degree items exist
    dense = np.zeros((n + m, n + m), dtype=np.float32)
    dense[row, col] = val
    assert np.allclose(dense, dense.T, rtol=3e-7, atol=0)  # bit-symmetric except mult=3 roundings
    assert (orc.degree_inv_sqrt(n, m, golden["train_user"], golden["train_item"]) == 0).sum() >= 4


def test_computer_matches_reference(golden):
    n, m, d, K, *_ = _g(golden)
    g = orc.sparse_graph(n, m, golden["train_user"], golden["train_item"])
    E0 = torch.from_numpy(golden["E0"])
    u, i = orc.computer(E0, g, K, n)
    assert np.array_equal(u.numpy(), golden["computer_users"])
    assert np.array_equal(i.numpy(), golden["computer_items"])
    folds = orc.sparse_graph(n, m, golden["train_user"], golden["train_item"], folds=7)
    u7, i7 = orc.computer(E0, folds, K, n)
    assert torch.equal(u7, u) and torch.equal(i7, i)  # A_split is not a numerics knob


def test_bpr_loss_grad_and_adam(golden):
    n, m, d, K, B, lr, decay = _g(golden)
    g = orc.sparse_graph(n, m, golden["train_user"], golden["train_item"])
    tu, tp, tn = (torch.from_numpy(golden[k]) for k in ("batch_users", "batch_pos", "batch_neg"))
    E = torch.from_numpy(golden["E0"]).clone().requires_grad_(True)
    loss, reg = orc.bpr_loss(E, g, K, n, tu, tp, tn)
    assert loss.item() == float(golden["loss"]) and reg.item() == float(golden["reg"])
    (loss + decay * reg).backward()
    assert np.allclose(E.grad.numpy(), golden["grad"], rtol=0, atol=1e-9)
    cf = orc.closed_form_grad(torch.from_numpy(golden["E0"]), g, K, n, tu, tp, tn, decay)
    assert (cf - E.grad).abs().max() / E.grad.abs().max() < 1e-5  # Horner identity
    om = orc.OracleModel(n, m, golden["train_user"], golden["train_item"], torch.from_numpy(golden["E0"]), K, lr, decay)
    l1 = om.stage_one(tu, tp, tn)
    assert np.allclose(om.weight.detach().numpy(), golden["E1"], rtol=0, atol=1e-7)
    l2 = om.stage_one(tu, tp, tn)
    assert np.allclose(om.weight.detach().numpy(), golden["E2"], rtol=0, atol=1e-7)
    assert abs(l1.item() - float(golden["step1_loss"])) < 1e-7
    assert abs(l2.item() - float(golden["step2_loss"])) < 1e-7


def test_sampler_decision_procedure_mt19937(golden, tiny_lists):
    train, _ = tiny_lists
    n, m, *_ = _g(golden)
    np.random.seed(123)
    S = orc.uniform_sample_mt(train, n, m, len(golden["train_user"]))
    assert np.array_equal(S, golden["sample_mt_seed123"])


def test_sampler_skips_empty_users():
    all_pos = [np.array([1, 2]), np.array([], dtype=np.int64), np.array([0])]
    S, valid = orc.uniform_sample_philox(all_pos, 3, 5, 200, seed=1, epoch=0)
    assert len(S) == int(valid.sum()) < 200
    assert not np.any(S[:, 0] == 1)
    for u, p, q in S:
        assert p in all_pos[u] and q not in all_pos[u]


def test_philox_known_answers():
    """Random123 v1.09 kat_vectors, philox4x32 with 10 rounds."""
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        got = orc.philox4x32_10(np.array([ctr], dtype=np.uint32), np.array([key], dtype=np.uint32))[0]
        assert [int(x) for x in got] == want


def test_philox_sampler_golden(golden, tiny_lists):
    train, _ = tiny_lists
    n, m, *_ = _g(golden)
    S, valid = orc.uniform_sample_philox(train, n, m, 2000, seed=2020, epoch=3)
    assert np.array_equal(S, golden["sample_philox_seed2020_epoch3"])
    assert valid.all()


def test_masked_topk_and_metrics(golden, tiny_lists):
    train, test = tiny_lists
    n, m, d, K, B, lr, decay = _g(golden)
    om = orc.OracleModel(n, m, golden["train_user"], golden["train_item"], torch.from_numpy(golden["E2"]), K, lr, decay)
    ks = [int(k) for k in golden["topks"]]
    users = list(test.keys())
    assert users == golden["eval_users"].tolist()
    rating = om.users_rating(torch.tensor(users))
    assert np.array_equal(rating[:8].numpy(), golden["raw_rating_first8"])
    vals, idx = orc.masked_topk(rating, [train[u] for u in users], max(ks))
    assert np.array_equal(idx.numpy(), golden["topk_raw_idx"])
    assert np.array_equal(vals.numpy(), golden["topk_raw_val"])
    res, _ = orc.evaluate(om, train, test, ks, 128)
    for mname in ("recall", "precision", "ndcg", "hr"):
        assert np.allclose(res[mname], golden[f"metric_raw_{mname}"], rtol=1e-12)


def test_masked_topk_tie_rule_and_mask_reentry():
    # ties -> lowest id first; masked items carry -1024 and CAN re-enter (trainer.py:137)
    r = torch.tensor([[1.0, 3.0, 3.0, 2.0, 3.0, 0.0], [-2000.0, -3000.0, 5.0, -1500.0, -4000.0, -2500.0]])
    vals, idx = orc.masked_topk(r, [np.array([], dtype=np.int64), np.array([2])], 3)
    assert idx[0].tolist() == [1, 2, 4]
    assert idx[1].tolist() == [2, 3, 0] and vals[1].tolist() == [-1024.0, -1500.0, -2000.0]


def test_philox_sampler_multi_negative_layout(tiny_lists, golden):
    """n_neg > 1: n_neg consecutive flat (u, pos, neg_t) rows per sample (lgcnssm.py:141 batches);
    sample i keeps the same user / positive as with n_neg == 1 and its first negative."""
    train, _ = tiny_lists
    n, m = int(golden["n_users"]), int(golden["m_items"])
    S1, _ = orc.uniform_sample_philox(train, n, m, 50, seed=7, epoch=1)
    S4, _ = orc.uniform_sample_philox(train, n, m, 50, seed=7, epoch=1, n_neg=4)
    assert S4.shape == (200, 3) and np.array_equal(S4[::4], S1)
    for u, p, q in S4:
        assert p in train[u] and q not in train[u]


def test_scipy_as_shipped_build_equals_the_golden_graph(golden):
    """The dok/lil route the reference ships (dataloader.py:226-244), which bench.py times, yields the
    live reference's graph bit for bit — so timing it is timing the reference's own build."""
    n, m = int(golden["n_users"]), int(golden["m_items"])
    A = orc.norm_adj_scipy_as_shipped(n, m, golden["train_user"], golden["train_item"]).tocoo()
    order = np.lexsort((A.col, A.row))
    assert np.array_equal(A.row[order], golden["adj_row"]) and np.array_equal(A.col[order], golden["adj_col"])
    assert np.array_equal(A.data[order].astype(np.float32), golden["adj_val"])
