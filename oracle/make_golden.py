"""Generate tests/golden/*.npz by running the LIVE reference (/root/reference) on CPU
and check the oracle restatement (oracle/lgcn_oracle.py) against it.

Run in the dev container only (`python oracle/make_golden.py`): the GPU box has
no /root/reference.  The shims are the ones listed in SURVEY.md §8c — none edits
the reference:
  1. sys.argv is set before `import world` (it parses at import, world.py:14) and
     world.device / config['device'] are pointed at the CPU (world.py:48-49);
  2. dataset.UserItemNet = csr_matrix(ones, (trainUser, trainItem)) restores the
     commented-out dataloader.py:164-165 so getSparseGraph() can build;
  3. model.device = 'cpu' on model.MF.LightGCN (MF.py:279 reads it);
  4. the data files Trainer / metric.py np.load at fixed relative paths are
     created as dummies in a temp cwd; WANDB_MODE=disabled;
     Trainer.checkpoint_save_path is redirected (trainer.py:222 hard-codes /home/...).
"""
from __future__ import annotations

import os
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
GOLD = REPO / "tests" / "golden"

N_USERS, M_ITEMS, DIM, LAYERS, BATCH = 300, 400, 32, 3, 64
LR, DECAY = 1e-3, 1e-4
TOPKS = [10, 20]


def make_tiny(seed=7):
    """300 x 400 interactions with the nasty cases: duplicate (u,i) pairs in a train
    line (multi-edges), items that only occur in test (zero train degree), the max
    item id only in test, unsorted file order, duplicate users inside a batch."""
    rng = np.random.default_rng(seed)
    pop = 1.0 / np.arange(1, M_ITEMS - 4 + 1) ** 0.8
    pop /= pop.sum()
    train, test = [], {}
    for u in range(N_USERS):
        deg = int(np.clip(rng.lognormal(2.2, 0.7), 5, 120))
        items = rng.choice(M_ITEMS - 4, size=min(deg, M_ITEMS - 4), replace=False, p=pop)
        n_tr = max(1, int(np.ceil(0.8 * len(items))))
        tr = list(items[:n_tr])
        if u % 17 == 0:  # multi-edge: repeat one item in the train line
            tr.append(tr[0])
        if u % 29 == 0:
            tr.extend([tr[1], tr[1]])
        rng.shuffle(tr)
        train.append([int(i) for i in tr])
        te = [int(i) for i in items[n_tr:]]
        if u % 50 == 3:
            te.append(M_ITEMS - 1 - ((u // 50) % 4))  # items never seen in train; max id only in test
        if te:
            test[u] = te
    return train, test


def write_files(root: Path, train, test, suffix="t"):
    d = root / "data" / "cf" / suffix
    d.mkdir(parents=True)
    with open(d / f"train{suffix}.txt", "w") as f:
        for u, its in enumerate(train):
            f.write(" ".join([str(u)] + [str(i) for i in its]) + "\n")
    with open(d / f"test{suffix}.txt", "w") as f:
        for u, its in test.items():
            f.write(" ".join([str(u)] + [str(i) for i in its]) + "\n")
    # Trainer / metric fixtures (trainer.py:47-48, metric.py:107,120, trainer.py:214)
    (root / "data" / suffix).mkdir(parents=True)
    (root / "data" / "cb" / suffix).mkdir(parents=True)
    (root / "data" / "result" / "lgn").mkdir(parents=True)
    names = np.array([f"p{i}" for i in range(M_ITEMS)], dtype=object)
    np.save(root / "data" / suffix / f"product_names{suffix}.npy", names, allow_pickle=True)
    cats = np.empty(M_ITEMS, dtype=object)
    for i in range(M_ITEMS):
        cats[i] = [i % 7, i % 3]
    np.save(root / "data" / "cb" / suffix / f"product_categories{suffix}.npy", cats, allow_pickle=True)
    np.save(root / "data" / "cb" / "product_categories.npy", cats, allow_pickle=True)
    np.save(root / "data" / "cf" / "product_occurance.npy", np.ones(M_ITEMS) * 3.0)


def main():
    import torch
    torch.manual_seed(2020)
    train, test = make_tiny()
    tmp = Path(tempfile.mkdtemp(prefix="lgcn_golden_"))
    write_files(tmp, train, test)
    os.chdir(tmp)
    os.environ["WANDB_MODE"] = "disabled"
    sys.argv = ["main.py", "--model", "lgn", "--recdim", str(DIM), "--layer", str(LAYERS), "--suffix", "t",
                "--bpr_batch", str(BATCH), "--lr", str(LR), "--decay", str(DECAY), "--testbatch", "128",
                "--topks", str(TOPKS), "--wandb", "golden", "--epochs", "1", "--a_fold", "7"]
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(REPO))
    import world  # noqa: E402  (reference)
    world.device = "cpu"
    world.config["device"] = "cpu"
    from scipy.sparse import csr_matrix
    import dataloader as ref_dataloader
    from model import MF as ref_MF
    import negative_sample as ref_ns
    import trainer as ref_trainer

    from oracle import lgcn_oracle as orc

    out = {}

    # ------------------------------------------------------------------ dataset + graph
    ds = ref_dataloader.Loader(world.config, path=str(tmp / "data" / "cf"))
    ds.UserItemNet = csr_matrix((np.ones(len(ds.trainUser)), (ds.trainUser, ds.trainItem)),
                                shape=(ds.n_user, ds.m_item))
    n, m = ds.n_users, ds.m_items
    assert (n, m) == (N_USERS, M_ITEMS), (n, m)
    out.update(n_users=n, m_items=m, train_user=ds.trainUser, train_item=ds.trainItem,
               test_user=ds.testUser, test_item=ds.testItem)

    folds = ds.getSparseGraph()          # as shipped: A_split=True (world.py:46), a_fold=7 here
    assert isinstance(folds, list) and len(folds) == 7
    ds.Graph = None
    ds.split = False
    os.remove(tmp / "data" / "cf" / "s_pre_adj_mat.npz")  # stale-cache hazard (dataloader.py:218-221)
    full = ds.getSparseGraph()
    ref_idx, ref_val = full.indices().numpy(), full.values().numpy()
    row, col, val = orc.norm_adj_coo(n, m, ds.trainUser, ds.trainItem)
    assert np.array_equal(ref_idx[0], row) and np.array_equal(ref_idx[1], col), "graph structure differs"
    assert np.array_equal(ref_val, val), "graph values are not bit-identical"
    ofolds = orc.sparse_graph(n, m, ds.trainUser, ds.trainItem, folds=7)
    for a, b in zip(folds, ofolds):
        assert a.shape == b.shape and torch.equal(a.indices(), b.indices()) and torch.equal(a.values(), b.values())
    out.update(adj_row=row, adj_col=col, adj_val=val)
    print(f"[graph] nnz={len(val)} bit-identical to the live reference (unsplit and 7 folds)")

    # ------------------------------------------------------------------ model (legacy torch.sparse path)
    cfg = dict(world.config)
    cfg.update(pretrain=0, dropout=0, keep_prob=0.6, A_split=False)
    model = ref_MF.LightGCN(cfg, ds)
    model.device = "cpu"
    E0 = torch.cat([model.embedding_user.weight, model.embedding_item.weight]).detach().clone()
    out["E0"] = E0.numpy()
    users_o, items_o = model.computer()
    out["computer_users"], out["computer_items"] = users_o.detach().numpy(), items_o.detach().numpy()
    g = orc.sparse_graph(n, m, ds.trainUser, ds.trainItem)
    ou, oi = orc.computer(E0, g, LAYERS, n)
    assert torch.equal(ou, users_o) and torch.equal(oi, items_o), "computer() differs"
    ou7, oi7 = orc.computer(E0, ofolds, LAYERS, n)
    assert torch.equal(ou7, users_o), "fold split is not bit-identical"
    print("[computer] oracle == live reference (bit-exact), folds == unsplit")

    # batch with duplicate users / items
    rng = np.random.default_rng(11)
    bu = rng.integers(0, n, BATCH)
    bu[5] = bu[6] = bu[7]
    bp = np.array([train[u][rng.integers(0, len(train[u]))] for u in bu])
    bn = rng.integers(0, m, BATCH)
    bn[9] = bn[10]
    tu, tp, tn = (torch.from_numpy(x).long() for x in (bu, bp, bn))
    out.update(batch_users=bu, batch_pos=bp, batch_neg=bn)

    loss, reg = model.bpr_loss(tu, tp, tn)
    total = loss + cfg["decay"] * reg
    model.optim.zero_grad()
    total.backward()
    grad = torch.cat([model.embedding_user.weight.grad, model.embedding_item.weight.grad]).clone()
    out.update(loss=loss.item(), reg=reg.item(), grad=grad.numpy())
    ol, orr = orc.bpr_loss(E0.clone().requires_grad_(True), g, LAYERS, n, tu, tp, tn)
    assert ol.item() == loss.item() and orr.item() == reg.item(), "bpr_loss differs"
    cf = orc.closed_form_grad(E0, g, LAYERS, n, tu, tp, tn, cfg["decay"])
    rel = (cf - grad).abs().max() / grad.abs().max()
    assert rel < 1e-5, rel
    print(f"[bpr] loss={loss.item():.6f} reg={reg.item():.6f}; closed-form grad vs autograd rel={rel:.2e}")

    # two Adam steps through the reference's own stageOne
    om = orc.OracleModel(n, m, ds.trainUser, ds.trainItem, E0, LAYERS, LR, DECAY)
    l1 = model.stageOne(tu, tp, tn)
    E1 = torch.cat([model.embedding_user.weight, model.embedding_item.weight]).detach().clone()
    l2 = model.stageOne(tu, tp, tn)
    E2 = torch.cat([model.embedding_user.weight, model.embedding_item.weight]).detach().clone()
    o1 = om.stage_one(tu, tp, tn)
    assert torch.allclose(om.weight.detach(), E1, rtol=0, atol=1e-7), (om.weight.detach() - E1).abs().max()
    o2 = om.stage_one(tu, tp, tn)
    assert torch.allclose(om.weight.detach(), E2, rtol=0, atol=1e-7)
    assert abs(o1.item() - l1.item()) < 1e-7 and abs(o2.item() - l2.item()) < 1e-7
    out.update(step1_loss=l1.item(), step2_loss=l2.item(), E1=E1.numpy(), E2=E2.numpy())
    print(f"[stageOne] two Adam steps match (loss {l1.item():.6f} -> {l2.item():.6f})")

    # ------------------------------------------------------------------ sampler (reference RNG)
    np.random.seed(123)
    S_ref = ref_ns.UniformSample(ds)
    np.random.seed(123)
    S_orc = orc.uniform_sample_mt(ds.allPos, n, m, ds.trainDataSize)
    assert np.array_equal(S_ref, S_orc), "sampler decision procedure differs"
    out["sample_mt_seed123"] = S_ref
    print(f"[sampler] {len(S_ref)} triples identical under numpy MT19937 seed 123")

    # ------------------------------------------------------------------ eval: Trainer.test on the E2 model
    ref_trainer.Trainer.checkpoint_save_path = staticmethod(lambda config: str(tmp / "ckpt.pth"))
    tr = ref_trainer.Trainer(world.config, ds, model)
    world.config["multicore"] = False
    res = tr.test()
    # the legacy class applies a sigmoid (MF.py:216); the registered class scores raw (lgcn.py:124)
    model.eval()
    eval_users = list(ds.testDict.keys())
    with torch.no_grad():
        au, ai = model.computer()
        raw = torch.matmul(au[torch.tensor(eval_users)], ai.t())
        sig = model.getUsersRating(torch.tensor(eval_users))
    out["eval_users"] = np.array(eval_users)
    out["raw_rating_first8"] = raw[:8].numpy()
    for name, rating in (("raw", raw), ("sig", sig)):
        vals, idx = orc.masked_topk(rating, [ds.allPos[u] for u in eval_users], max(TOPKS))
        rr = rating.clone()
        ex_r, ex_c = [], []
        for r_i, its in enumerate(ds.getUserPosItems(eval_users)):
            ex_r.extend([r_i] * len(its)); ex_c.extend(its)
        rr[ex_r, ex_c] = -(1 << 10)
        tv, ti = torch.topk(rr, k=max(TOPKS))       # trainer.py:137-138
        assert torch.equal(tv, vals), name
        ties = (tv[:, 1:] == tv[:, :-1]).any().item()
        if not ties:
            assert torch.equal(ti, idx), name
        out[f"topk_{name}_idx"], out[f"topk_{name}_val"] = idx.numpy(), vals.numpy()
        bm = orc.batch_metrics([ds.testDict[u] for u in eval_users], idx.numpy(), TOPKS)
        for k_, v_ in bm.items():
            out[f"metric_{name}_{k_}"] = v_ / len(eval_users)
        print(f"[eval/{name}] ties_in_topk={ties} recall@{TOPKS}={bm['recall'] / len(eval_users)}")
    for mname in ("recall", "precision", "ndcg", "hr"):
        assert np.allclose(res[mname], out[f"metric_sig_{mname}"], rtol=1e-12, atol=0), (mname, res[mname])
    om.weight.data.copy_(E2)
    ores, _ = orc.evaluate(om, ds.allPos, ds.testDict, TOPKS, 128)
    for mname in ("recall", "precision", "ndcg", "hr"):
        assert np.allclose(ores[mname], out[f"metric_raw_{mname}"], rtol=1e-12), mname
    print("[metrics] oracle == Trainer.test() of the live reference:", {k: res[k] for k in ("recall", "ndcg")})

    # ------------------------------------------------------------------ Philox sampler KAT (our spec)
    S_ph, valid = orc.uniform_sample_philox(ds.allPos, n, m, 2000, seed=2020, epoch=3)
    out["sample_philox_seed2020_epoch3"] = S_ph
    out["config"] = np.array([DIM, LAYERS, BATCH], dtype=np.int64)
    out["hyper"] = np.array([LR, DECAY], dtype=np.float64)
    out["topks"] = np.array(TOPKS, dtype=np.int64)

    GOLD.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLD / "tiny_ref.npz", **out)
    print("wrote", GOLD / "tiny_ref.npz", f"{(GOLD / 'tiny_ref.npz').stat().st_size / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
