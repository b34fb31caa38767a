"""Golden vectors for the SURVEY §8(f-4) variants, from the LIVE reference classes.

Run in the dev container only (`python oracle/make_golden_variants.py`, after
`oracle/make_golden.py`): imports `/root/reference/model/{radj,lgcn}.py` and the capped
`UniformSample` of `/root/reference/ddp_lgcn.py` unmodified, on the tiny data set frozen in
tests/golden/tiny_ref.npz, checks the oracle restatements against them and writes
tests/golden/variants_ref.npz.

Those modules import two third-party packages that are neither vendored nor listed in the
reference's requirements.txt (version unpinned) and are absent here: `torch_geometric`
(LGConv, NeighborSampler) and `torch_scatter` (scatter).  They are replaced in `sys.modules` by
stubs carrying the published semantics (`oracle.lgcn_oracle.lgconv_pyg`, `scatter_sum`), so
everything the reference itself wrote (degree tables, edge divisors, layer loop, loss, Adam step,
sampler loop) is pinned bit for bit and only the two library calls remain "parity unpinned".
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
GOLD = REPO / "tests" / "golden"
R_EXP = 0.3          # rAdjGCN exponent (parse.py:49 default 0.5 would hide a src/dst swap)
CAP = 3              # POSITIVE_NUM_LIMIT scaled to the tiny data (ddp_lgcn.py:34 ships 3000)


def install_stubs(orc):
    import torch

    class LGConv(torch.nn.Module):
        def forward(self, x, edge_index):
            return orc.lgconv_pyg(x, edge_index)

    def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
        assert dim == 0 and out is not None and reduce == "sum"
        return orc.scatter_sum(src, index, out)

    tg = types.ModuleType("torch_geometric")
    tg_nn = types.ModuleType("torch_geometric.nn")
    tg_conv = types.ModuleType("torch_geometric.nn.conv")
    tg_loader = types.ModuleType("torch_geometric.loader")
    tg_conv.LGConv = LGConv
    tg_nn.conv = tg_conv
    tg.nn = tg_nn
    tg_loader.NeighborSampler = object
    tg.loader = tg_loader
    ts = types.ModuleType("torch_scatter")
    ts_sc = types.ModuleType("torch_scatter.scatter")
    ts_sc.scatter = scatter
    ts.scatter = ts_sc
    sys.modules.update({"torch_geometric": tg, "torch_geometric.nn": tg_nn, "torch_geometric.nn.conv": tg_conv,
                        "torch_geometric.loader": tg_loader, "torch_scatter": ts, "torch_scatter.scatter": ts_sc})


class TinyDataset:
    """The attributes the reference model classes and samplers read (model/lgcn.py:49-54,
    ddp_lgcn.py:549-552)."""

    def __init__(self, g):
        self.n_users, self.m_items = int(g["n_users"]), int(g["m_items"])
        self.trainUser, self.trainItem = g["train_user"], g["train_item"]
        self.trainDataSize = len(self.trainUser)
        order = np.argsort(self.trainUser, kind="stable")
        cnt = np.bincount(self.trainUser, minlength=self.n_users)
        ptr = np.concatenate([[0], np.cumsum(cnt)])
        items = self.trainItem[order]
        self.allPos = [items[ptr[u]:ptr[u + 1]] for u in range(self.n_users)]
        import torch
        self.item_oc = torch.from_numpy(np.bincount(self.trainItem, minlength=self.m_items).astype(np.float32) + 1)


def main():
    import torch
    g = dict(np.load(GOLD / "tiny_ref.npz"))
    d, K, B = (int(x) for x in g["config"])
    lr, decay = (float(x) for x in g["hyper"])
    tmp = Path(tempfile.mkdtemp(prefix="lgcn_golden_var_"))
    os.chdir(tmp)
    os.environ["WANDB_MODE"] = "disabled"
    sys.argv = ["main.py", "--model", "lgn", "--recdim", str(d), "--layer", str(K), "--bpr_batch", str(B),
                "--lr", str(lr), "--decay", str(decay), "--r", str(R_EXP)]
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(REPO))
    from oracle import lgcn_oracle as orc
    install_stubs(orc)
    import world  # noqa: E402 (reference)
    world.device = "cpu"
    world.config["device"] = "cpu"
    cfg = dict(world.config)
    ds = TinyDataset(g)
    n, m = ds.n_users, ds.m_items
    E0 = torch.from_numpy(g["E0"])
    tu, tp, tn = (torch.from_numpy(g[k]).long() for k in ("batch_users", "batch_pos", "batch_neg"))
    edge = orc.directed_edges(n, ds.trainUser, ds.trainItem)
    out = {"r": np.float64(R_EXP), "cap": np.int64(CAP)}

    def run_model(cls, tag, fwd):
        model = cls(cfg, ds)
        with torch.no_grad():
            model.all_embedding.weight.copy_(E0)
        u, i = model.forward()
        ou, oi = fwd(E0)
        assert torch.equal(u, ou) and torch.equal(i, oi), f"{tag}: forward differs"
        loss, reg = model.bpr_loss(tu, tp, tn)
        model.optim.zero_grad()
        (loss + cfg["decay"] * reg).backward()
        grad = model.all_embedding.weight.grad.clone()
        w = E0.clone().requires_grad_(True)
        fu, fi = fwd(w)
        ol, orr = orc.bpr_loss_from(fu, fi, w, n, tu, tp, tn)
        assert ol.item() == loss.item() and orr.item() == reg.item(), f"{tag}: loss differs"
        (ol + cfg["decay"] * orr).backward()
        assert torch.allclose(w.grad, grad, rtol=0, atol=1e-9), f"{tag}: grad differs"
        l1 = model.stageOne(tu, tp, tn).item()
        E1 = model.all_embedding.weight.detach().clone()
        l2 = model.stageOne(tu, tp, tn).item()
        E2 = model.all_embedding.weight.detach().clone()
        out.update({f"{tag}_users": u.detach().numpy(), f"{tag}_items": i.detach().numpy(),
                    f"{tag}_loss": loss.item(), f"{tag}_reg": reg.item(), f"{tag}_grad": grad.numpy(),
                    f"{tag}_step1_loss": l1, f"{tag}_step2_loss": l2, f"{tag}_E1": E1.numpy(), f"{tag}_E2": E2.numpy()})
        print(f"[{tag}] forward/loss bit-exact vs oracle; loss={loss.item():.6f} -> {l1:.6f} -> {l2:.6f}")

    from model import radj as ref_radj
    run_model(ref_radj.rAdjGCN, "radj", lambda w: orc.radj_forward(w, edge, K, n, R_EXP))

    from model import lgcn as ref_lgcn
    run_model(ref_lgcn.LightGCN, "pyg", lambda w: orc.lgconv_forward(w, edge, K, n))
    # the PyG form and the torch.sparse form (model/MF.py) agree to fp32 rounding
    rel = np.abs(out["pyg_users"] - g["computer_users"]).max() / np.abs(g["computer_users"]).max()
    assert rel < 1e-6, rel
    print(f"[pyg] LGConv form vs torch.sparse golden: rel {rel:.2e}")

    # RGCN (model/rgcn.py) reads a private csv at an absolute path (:60), so the class cannot be
    # constructed; its forward is the LightGCN loop on purchase+favourite edges (:108-116): freeze
    # the oracle's output for a synthetic favourite list (parity unpinned for this one).
    rng = np.random.default_rng(5)
    fav_u = rng.integers(0, n, 500)
    fav_i = rng.integers(0, m, 500)
    edge_f = orc.directed_edges(n, ds.trainUser, ds.trainItem, fav_u, fav_i)
    ru, ri = orc.lgconv_forward(E0, edge_f, K, n)
    out.update(fav_user=fav_u, fav_item=fav_i, rgcn_users=ru.numpy(), rgcn_items=ri.numpy())

    # capped sampler of the DDP script (ddp_lgcn.py:541-582)
    import ddp_lgcn as ref_ddp
    ref_ddp.POSITIVE_NUM_LIMIT = CAP
    ref_ddp.tqdm = lambda it, *a, **k: it
    np.random.seed(321)
    S_ref = ref_ddp.UniformSample(ds)
    np.random.seed(321)
    S_orc = orc.capped_sample_mt(ds.allPos, m, ds.trainDataSize * ref_ddp.TRAIN_ITERATIVE, CAP)
    assert np.array_equal(S_ref, S_orc), "capped sampler decision procedure differs"
    assert np.bincount(S_ref[:, 1]).max() == CAP
    out["capped_mt_seed321"] = S_ref
    out["capped_philox_seed9_epoch1"] = orc.capped_sample_philox(ds.allPos, n, m, 3000, seed=9, epoch=1, limit=CAP)
    print(f"[capped sampler] {len(S_ref)} of {ds.trainDataSize * 3} triples kept, identical under MT19937 seed 321")

    # edge dropout of the legacy class (model/MF.py:158-192): same torch seed -> same mask
    from model import MF as ref_MF
    KEEP = 0.6
    ds.getSparseGraph = lambda: orc.sparse_graph(n, m, ds.trainUser, ds.trainItem)   # == the reference's (make_golden.py)
    cfg_d = dict(cfg)
    cfg_d.update(latent_dim_rec=d, lightGCN_n_layers=K, keep_prob=KEEP, A_split=False, pretrain=0, dropout=1)
    mdl = ref_MF.LightGCN(cfg_d, ds)
    mdl.device = "cpu"
    with torch.no_grad():
        mdl.embedding_user.weight.copy_(E0[:n])
        mdl.embedding_item.weight.copy_(E0[n:])
    mdl.train()
    import contextlib, io
    torch.manual_seed(99)
    with contextlib.redirect_stdout(io.StringIO()):      # the reference prints "droping" per call
        loss_d, reg_d = mdl.bpr_loss(tu, tp, tn)
    mdl.optim.zero_grad()
    (loss_d + cfg["decay"] * reg_d).backward()
    grad_d = torch.cat([mdl.embedding_user.weight.grad, mdl.embedding_item.weight.grad]).clone()
    g_full = ds.getSparseGraph()
    torch.manual_seed(99)
    mask = (torch.rand(len(g_full.values())) + KEEP).int().bool()
    gd = orc.dropout_graph(g_full, mask, KEEP)
    w = E0.clone().requires_grad_(True)
    ol, orr = orc.bpr_loss(w, gd, K, n, tu, tp, tn)
    assert ol.item() == loss_d.item() and orr.item() == reg_d.item(), "dropout: loss differs"
    (ol + cfg["decay"] * orr).backward()
    assert torch.allclose(w.grad, grad_d, rtol=0, atol=1e-9), "dropout: grad differs"
    du, di = orc.computer(E0, gd, K, n)
    out.update(dropout_keep=np.float64(KEEP), dropout_mask=mask.numpy(), dropout_users=du.numpy(), dropout_items=di.numpy(),
               dropout_loss=loss_d.item(), dropout_reg=reg_d.item(), dropout_grad=grad_d.numpy())
    assert 0.5 < mask.float().mean() < 0.7 and not torch.equal(gd.to_dense(), gd.to_dense().t())
    print(f"[dropout] keep {mask.float().mean():.3f} of {len(mask)} entries (asymmetric); loss/grad match the live reference")

    # popularity-weighted positive pick of UniformSampling (negative_sample.py:12-69): the class
    # unpickles per-user probabilities from a fixed relative path (:22-36); feed it ours.
    import pickle
    import negative_sample as ref_ns
    pop = np.bincount(ds.trainItem, minlength=m).astype(np.float64)
    probs = [(pop[p] ** -0.5) / (pop[p] ** -0.5).sum() if len(p) else np.zeros(0) for p in ds.allPos]
    (tmp / "data" / "sample_prob").mkdir(parents=True)
    with open(tmp / "data" / "sample_prob" / "sample_prob_05.pkl", "wb") as f:
        pickle.dump(probs, f)
    ds.n_user = n
    ref_ns.tqdm = lambda it, *a, **k: it
    sampler = ref_ns.UniformSampling(ds, {"sample_pow": 0.5})
    np.random.seed(77)
    users_drawn = np.random.randint(0, n, 4000)
    ret = {}
    sampler.sample_parallel(0, 1, users_drawn, ds.allPos, ret)
    np.random.seed(77)
    W_orc = orc.weighted_sample_mt(ds.allPos, probs, n, m, 4000)
    assert np.array_equal(ret[0], W_orc), "weighted sampler decision procedure differs"
    cdfs = [orc.normalised_cdf(p) if len(p) else np.zeros(0, np.float32) for p in probs]
    out["weighted_mt_seed77"] = ret[0]
    out["weighted_probs_flat"] = np.concatenate(probs)
    out["weighted_philox_seed4_epoch2"] = orc.uniform_sample_philox(ds.allPos, n, m, 3000, seed=4, epoch=2, pos_cdf=cdfs)[0]
    print(f"[weighted sampler] {len(W_orc)} triples identical under MT19937 seed 77 (sample_pow=0.5)")

    np.savez_compressed(GOLD / "variants_ref.npz", **out)
    print("wrote", GOLD / "variants_ref.npz", f"{(GOLD / 'variants_ref.npz').stat().st_size / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
