"""Row-partition logic and the per-layer all-gather orchestration on CPU: world_size 2,
gloo backend, with the local SpMM injected (the CUDA kernel is the only thing replaced)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from furusato_recommend_b200 import graph as G
from furusato_recommend_b200.parallel import DistPropagator, RowPartition, exchange_rows
from oracle import lgcn_oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden", "tiny_ref.npz")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cpu_local_spmm(rp, col, dinv):
    """lgcn_propagate_layer's contract (scale_src=0) in plain torch, for the rank's local CSR
    (CSR rows = the rank's real rows; buffers are padded to R = len(dinv))."""
    n_real, R = rp.numel() - 1, dinv.numel()
    rows = torch.repeat_interleave(torch.arange(n_real), rp[1:] - rp[:-1])

    def f(src_full, dst=None, base=None, acc_in=None, acc_out=None, acc_scale=1.0, **kw):
        s = torch.zeros((R, src_full.shape[1])).index_add_(0, rows, src_full[col.long()].float())
        x = dinv[:, None] * s
        t = base + x if base is not None else x
        if dst is not None:
            dst.copy_(dinv[:, None] * t)
        if acc_out is not None:
            acc_out.copy_((acc_in + x) * acc_scale)
        if kw.get("grad_mode") == 1:
            kw["grad"].copy_(t * kw["inv_layers"] + kw["reg_coef"] * kw["cnt"][:, None].float() * kw["emb"])
    return f


def _worker(rank, world, port, ret, side_split=False):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = dict(np.load(GOLD))
        n, m = int(g["n_users"]), int(g["m_items"])
        K = int(g["config"][1])
        csr = G.build_csr_graph(n, m, torch.from_numpy(g["train_user"]), torch.from_numpy(g["train_item"]))
        part = RowPartition(csr.rowptr, world, n_users=n if os.environ.get("LGCN_TEST_TWO_SIDED", "1") == "1" else None,
                            side_split=side_split)
        if side_split:  # rank 0 owns users only, rank 1 items only; each pushes to the other side only
            assert part.side_split and part.readers_of(rank) == [1 - rank]
            assert part.ranges(0)[1][0] == part.ranges(0)[1][1] and part.ranges(1)[0][0] == part.ranges(1)[0][1]
        rp, colp, dl = part.local_csr(rank, csr.rowptr, csr.col, csr.dinv)
        prop = DistPropagator(part, rank, dl, K, _cpu_local_spmm(rp, colp, dl))
        E = torch.from_numpy(g["E0"])
        emb = part.shard(rank, E)
        acc, out = torch.empty_like(emb), torch.empty_like(emb)
        prop.forward(emb, acc, out)
        full = torch.empty((world * part.R, E.shape[1]))
        dist.all_gather_into_tensor(full, out)
        light = part.unshard(full)
        ref = torch.from_numpy(np.concatenate([g["computer_users"], g["computer_items"]]))
        err_f = float((light - ref).abs().max() / ref.abs().max())

        # backward: Horner on a dense seed == autograd through the oracle's computer()
        gen = torch.Generator().manual_seed(5)
        W = torch.randn(E.shape, generator=gen)
        Ew = E.clone().requires_grad_(True)
        sg = orc.sparse_graph(n, m, g["train_user"], g["train_item"])
        ou, oi = orc.computer(Ew, sg, K, n)
        (torch.cat([ou, oi]) * W).sum().backward()
        Gl = part.shard(rank, W)
        grad = torch.empty_like(Gl)
        prop.backward(Gl, grad_mode=1, inv_layers=1.0 / (K + 1), reg_coef=0.0,
                      cnt=torch.zeros(part.R, dtype=torch.int32), emb=emb, grad=grad)
        dist.all_gather_into_tensor(full, grad)
        err_b = float((part.unshard(full) - Ew.grad).abs().max() / Ew.grad.abs().max())

        # row exchange: arbitrary global rows, duplicates included
        ids = part.to_padded(torch.tensor([0, n + 3, 5, n + m - 1, 5, 299, n]))
        rows, mine = exchange_rows(part, rank, emb, ids)
        err_x = float((rows - E[torch.tensor([0, n + 3, 5, n + m - 1, 5, 299, n])]).abs().max())
        # sparse seed (the BPR step's <= 3B rows, duplicates included): building the first layer's
        # gathered source locally (g0) == exchanging dinv (.) G
        gids = torch.tensor([0, n + 3, 5, n + m - 1, 5, 299, n])
        G_c = torch.randn((len(gids), E.shape[1]), generator=gen)
        pad = part.to_padded(gids)
        own = (pad // part.R) == rank
        Gs = torch.zeros_like(emb).index_add_(0, pad % part.R, torch.where(own[:, None], G_c, torch.zeros(())))
        dinv_pad = torch.zeros(world * part.R)
        dinv_pad[part.to_padded(torch.arange(n + m))] = csr.dinv
        ga, gb = torch.empty_like(Gs), torch.empty_like(Gs)
        kw = dict(grad_mode=1, inv_layers=1.0 / (K + 1), reg_coef=0.0, cnt=torch.zeros(part.R, dtype=torch.int32), emb=emb)
        prop.backward(Gs, grad=ga, **kw)
        prop.backward(Gs, g0=(dinv_pad, pad, G_c), grad=gb, **kw)
        err_g0 = float((ga - gb).abs().max() / ga.abs().max())
        ret[rank] = (err_f, err_b, err_x, int(mine.sum()), part.R, [c.tolist() for c in part.cuts], err_g0)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_users", [None, 3])
def test_row_partition_indexing(n_users):
    rowptr = torch.tensor([0, 4, 4, 10, 11, 11, 30, 31, 40])
    for world in (1, 2, 3, 4):
        p = RowPartition(rowptr, world, n_users=n_users)
        assert sum(p.rows) == 8 and p.R == max(p.rows)
        ids = torch.arange(8)
        pad = p.to_padded(ids)
        assert len(torch.unique(pad)) == 8 and int(pad.max()) < world * p.R
        own = p.owner(ids)
        assert torch.equal(own, pad // p.R)
        for r in range(world):
            for lo, hi in p.ranges(r):
                assert bool((own[lo:hi] == r).all())
        x = torch.arange(16.0).reshape(8, 2)
        gathered = torch.cat([p.shard(r, x) for r in range(world)])
        assert torch.equal(p.unshard(gathered), x)
        assert torch.equal(gathered[pad], x)          # padded id addresses the gathered layout
        # every edge lands in exactly one local CSR with a remapped column
        col = torch.arange(40, dtype=torch.int32) % 8
        tot = 0
        for r in range(world):
            rp, colp, dl = p.local_csr(r, rowptr, col, torch.ones(8))
            assert rp.numel() == p.rows[r] + 1 and int(rp[-1]) == colp.numel() and dl.numel() == p.R
            tot += colp.numel()
        assert tot == 40
    if n_users is not None:  # side split: users on the first W/2 ranks, items on the rest
        for world in (2, 4):
            p = RowPartition(rowptr, world, n_users=n_users, side_split=True)
            assert p.side_split and sum(p.rows) == 8
            ids = torch.arange(8)
            pad, own = p.to_padded(ids), p.owner(ids)
            assert len(torch.unique(pad)) == 8 and torch.equal(own, pad // p.R)
            assert bool((own[:n_users] < world // 2).all()) and bool((own[n_users:] >= world // 2).all())
            x = torch.arange(16.0).reshape(8, 2)
            gathered = torch.cat([p.shard(r, x) for r in range(world)])
            assert torch.equal(p.unshard(gathered), x) and torch.equal(gathered[pad], x)
            assert p.readers_of(0) == list(range(world // 2, world)) and p.readers_of(world - 1) == list(range(world // 2))
            col = torch.arange(40, dtype=torch.int32) % 8
            assert sum(p.local_csr(r, rowptr, col, torch.ones(8))[1].numel() for r in range(world)) == 40
        assert not RowPartition(rowptr, 3, n_users=n_users, side_split=True).side_split   # odd world: falls back
    if n_users is not None:  # two-sided: both sides are cut separately, rows stay balanced
        p = RowPartition(rowptr, 2, n_users=n_users)
        assert len(p.cuts) == 2 and p.cuts[0][0] == 0 and p.cuts[0][-1] == 3 and p.cuts[1][0] == 3 and p.cuts[1][-1] == 8


@pytest.mark.timeout(300)
@pytest.mark.parametrize("side_split", [False, True])
def test_partitioned_propagation_world2_gloo(side_split):
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret, side_split), nprocs=world, join=True)
        assert len(ret) == world
        for rank in range(world):
            err_f, err_b, err_x, n_mine, R, starts, err_g0 = ret[rank]
            assert err_f < 1e-5, f"forward mismatch on rank {rank}: {err_f}"
            assert err_b < 1e-5, f"backward mismatch on rank {rank}: {err_b}"
            assert err_x == 0.0
            assert err_g0 < 1e-6, f"local layer-0 source differs from the exchanged one on rank {rank}: {err_g0}"
        assert ret[0][3] + ret[1][3] == 7 and ret[0][5] == ret[1][5]


@pytest.mark.parametrize("world", [2, 3])
def test_reduce_partition_graphs_reproduce_the_propagation(world):
    """Index logic of the "reduce" partition (parallel.reduce_local_graphs), single process: user rows gathered
    from the padded item table (graph A) plus, for every item, the sum over ranks of the partial sums over each
    rank's local users (graph B, owner-block numbering) must equal A_hat X of the oracle."""
    from furusato_recommend_b200.parallel import reduce_local_graphs
    g = dict(np.load(GOLD))
    n, m = int(g["n_users"]), int(g["m_items"])
    csr = G.build_csr_graph(n, m, torch.from_numpy(g["train_user"]), torch.from_numpy(g["train_item"]))
    part = RowPartition(csr.rowptr, world, n_users=n)
    N, R = n + m, part.R
    X = torch.from_numpy(g["E0"]).double()
    graph = orc.sparse_graph(n, m, g["train_user"], g["train_item"]).to_dense().double()
    want = graph @ X
    Z = csr.dinv.double()[:, None] * X                                   # pre-scaled activations
    padded = part.to_padded(torch.arange(N))
    Zfull = torch.zeros(world * R, X.shape[1], dtype=torch.float64)
    Zfull[padded] = Z
    got = torch.zeros_like(want)
    partial_sum = None
    R_i = None
    for r in range(world):
        rp_a, col_a, rp_b, col_b, R_i, sizes = reduce_local_graphs(part, r, csr.rowptr, csr.col)
        (u_lo, u_hi), _ = part.ranges(r)
        rows_a = torch.repeat_interleave(torch.arange(u_hi - u_lo), rp_a[1:] - rp_a[:-1])
        s = torch.zeros(u_hi - u_lo, X.shape[1], dtype=torch.float64).index_add_(0, rows_a, Zfull[col_a.long()])
        got[u_lo:u_hi] = csr.dinv.double()[u_lo:u_hi, None] * s       # user rows: local gather from the item table
        rows_b = torch.repeat_interleave(torch.arange(world * R_i), rp_b[1:] - rp_b[:-1])
        p = torch.zeros(world * R_i, X.shape[1], dtype=torch.float64).index_add_(0, rows_b, Z[u_lo:u_hi][col_b.long()])
        partial_sum = p if partial_sum is None else partial_sum + p       # what the owners' reduce kernel adds up
        assert int(col_b.max()) < u_hi - u_lo and int(rp_b[-1]) == col_b.numel()
    item_cuts = part.cuts[1]
    for b in range(world):
        lo, hi = int(item_cuts[b]), int(item_cuts[b + 1])
        got[lo:hi] = csr.dinv.double()[lo:hi, None] * partial_sum[b * R_i:b * R_i + (hi - lo)]
        assert float(partial_sum[b * R_i + (hi - lo):(b + 1) * R_i].abs().sum()) == 0.0   # padding rows stay empty
    assert float((got - want).abs().max()) < 1e-6 * float(want.abs().max())


def _worker_shared_graph(rank, world, port, ret):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    import bench
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        W = dict(n_users=1500, m_items=400, n_interactions=40_000, seed=9)
        n, m, tu, ti, su, si = bench.shared_device_graph(W, "cpu", rank, world)
        ret[rank] = (n, m, tu.clone(), ti.clone(), su.clone(), si.clone())
    finally:
        dist.destroy_process_group()


def test_big_graphs_are_generated_once_and_broadcast():
    """bench.shared_device_graph: rank 0 generates, every rank receives the same tensors (gloo, world 2) — the
    ranks of a row-partitioned model must cut their rows out of ONE graph."""
    from furusato_recommend_b200.synthetic import bipartite
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker_shared_graph, args=(world, port, ret), nprocs=world, join=True)
        ref = bipartite(1500, 400, 40_000, seed=9)
        for rank in range(world):
            got = ret[rank]
            assert got[0] == ref[0] and got[1] == ref[1]
            for a, b in zip(got[2:], ref[2:]):
                assert torch.equal(a, b)
