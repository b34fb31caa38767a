"""Row-partitioned training on 2 GPUs (NCCL) against the single-GPU path.  Needs >= 2 devices:
run with `gpurun --gpus 2`; skipped on the 1-GPU round-end box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "tiny_ref.npz")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret, exchange="auto", partition="side_split"):
    import torch.distributed as dist
    from furusato_recommend_b200 import LightGCN
    from furusato_recommend_b200.dataloader import BasicDataset
    from furusato_recommend_b200.parallel import DistLightGCN
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    try:
        g = dict(np.load(GOLD))
        d, K, B = (int(x) for x in g["config"])
        lr, decay = (float(x) for x in g["hyper"])
        cfg = dict(recdim=d, layer=K, lr=lr, decay=decay, bpr_batch_size=B, device=dev, test_u_batch_size=128,
                   dist_exchange=exchange, dist_partition=partition)
        ds = BasicDataset(int(g["n_users"]), int(g["m_items"]), g["train_user"], g["train_item"], g["test_user"],
                          g["test_item"], config=cfg, device=dev)
        E0 = torch.from_numpy(g["E0"]).to(dev)
        dm = DistLightGCN(cfg, ds, rank, world)
        dm.load_global_embedding(E0)
        u, p, q = (torch.from_numpy(g[k]).to(dev) for k in ("batch_users", "batch_pos", "batch_neg"))
        out = dm.computer_local()
        light = dm.part.unshard(_gather(dm, out))
        ref = torch.from_numpy(np.concatenate([g["computer_users"], g["computer_items"]])).to(dev)
        e_prop = float((light - ref).abs().max() / ref.abs().max())
        l1 = float(dm.fused_step(u, p, q))
        l2 = float(dm.fused_step(u, p, q))
        E2 = dm.gather_embedding()
        e_emb = float((E2.cpu() - torch.from_numpy(g["E2"])).abs().max())
        users = torch.from_numpy(g["eval_users"]).to(dev)
        dm.computer_local()
        mine, idx, val = dm.topk_user_shard(users, 20, precision="fp32")
        sm = LightGCN(cfg, ds)
        with torch.no_grad():
            sm.all_embedding.weight.copy_(E2)
        sm.eval()
        sidx, sval = sm.getUsersTopK(mine, 20, precision="fp32")
        same = float((sidx == idx).float().mean())
        ret[rank] = (e_prop, l1, l2, e_emb, same, float(g["step1_loss"]), float(g["step2_loss"]), dm.exchange)
    finally:
        dist.destroy_process_group()


def _gather(dm, local):
    import torch.distributed as dist
    full = torch.empty((dm.world * dm.part.R, dm.d), dtype=torch.float32, device=dm.device)
    dist.all_gather_into_tensor(full, local)
    return full


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("exchange,partition", [("nccl", "two_sided"), ("push", "two_sided"), ("push", "side_split"),
                                                ("nccl", "side_split"), ("push", "reduce")])
@pytest.mark.timeout(120)
def test_row_partitioned_training_matches_reference_2gpu(exchange, partition):
    """nccl: ncclAllGather per layer; push: all-gather fused into the SpMM epilogue (NVLink peer stores) and
    the whole step replayed as one CUDA graph.  side_split: users on rank 0, items on rank 1.  reduce: user rows
    never travel — per-item partial sums go to the item's owner (lgcn_reduce_rows), only item rows are gathered."""
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret, exchange, partition), nprocs=world, join=True)
        for rank in range(world):
            e_prop, l1, l2, e_emb, same, g1, g2, used = ret[rank]
            assert used == exchange
            assert e_prop < 1e-5, e_prop
            assert abs(l1 - g1) < 1e-5 * g1 and abs(l2 - g2) < 1e-5 * g2, (l1, g1, l2, g2)
            assert e_emb < 1e-6, e_emb
            assert same > 0.999, same
        assert ret[0][1] == ret[1][1]  # the loss is identical on every rank


def _worker_partial(rank, world, port, ret, exchange):
    """full, partial, full, full batches: the captured step graph must survive the eager partial batch
    (grow-only scratch), and an eval propagation between steps (the pre-scaled table is re-pushed)."""
    import torch.distributed as dist
    from furusato_recommend_b200 import LightGCN
    from furusato_recommend_b200.dataloader import BasicDataset
    from furusato_recommend_b200.parallel import DistLightGCN
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    try:
        g = dict(np.load(GOLD))
        d, K, B = (int(x) for x in g["config"])
        lr, decay = (float(x) for x in g["hyper"])
        cfg = dict(recdim=d, layer=K, lr=lr, decay=decay, bpr_batch_size=B, device=dev, test_u_batch_size=128,
                   dist_exchange=exchange)
        ds = BasicDataset(int(g["n_users"]), int(g["m_items"]), g["train_user"], g["train_item"], g["test_user"],
                          g["test_item"], config=cfg, device=dev)
        E0 = torch.from_numpy(g["E0"]).to(dev)
        u, p, q = (torch.from_numpy(g[k]).to(dev) for k in ("batch_users", "batch_pos", "batch_neg"))
        assert u.numel() == B
        sm = LightGCN(cfg, ds)
        with torch.no_grad():
            sm.all_embedding.weight.copy_(E0)
        sm.train()
        dm = DistLightGCN(cfg, ds, rank, world)
        dm.load_global_embedding(E0)
        h = B // 3
        losses, ref = [], []
        for step, sl in enumerate((slice(0, B), slice(0, h), slice(0, B), slice(0, B))):
            if step == 3:                      # an eval-only propagation between two replays
                dm.computer_local()
            losses.append(float(dm.fused_step(u[sl], p[sl], q[sl])))
            ref.append(float(sm.stageOne(u[sl], p[sl], q[sl])))
        dm.check_ids()
        e_emb = float((dm.gather_embedding() - sm.all_embedding.weight.detach()).abs().max())
        ret[rank] = (losses, ref, e_emb, dm._graph is not None)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("exchange", ["push", "nccl"])
@pytest.mark.timeout(120)
def test_partial_batch_between_graph_replays_2gpu(exchange):
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker_partial, args=(world, port, ret, exchange), nprocs=world, join=True)
        for rank in range(world):
            losses, ref, e_emb, captured = ret[rank]
            assert captured == (exchange == "push")
            for a, b in zip(losses, ref):
                assert abs(a - b) < 1e-5 * abs(b), (losses, ref)
            assert e_emb < 1e-6, e_emb


def _worker_user_heavy(rank, world, port, ret):
    """dist_partition='auto' on a user-heavy graph (6000 users x 500 items): the reduce partition is picked
    (parallel.DistLightGCN, measured on cfg-3) and two steps agree with the single-GPU model."""
    import torch.distributed as dist
    from furusato_recommend_b200 import LightGCN
    from furusato_recommend_b200.dataloader import BasicDataset
    from furusato_recommend_b200.parallel import DistLightGCN
    from furusato_recommend_b200.synthetic import bipartite
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    try:
        n, m, tu, ti, su, si = bipartite(6000, 500, 120_000, seed=5)
        B = 512
        cfg = dict(recdim=64, layer=3, lr=1e-3, decay=1e-4, bpr_batch_size=B, device=dev, test_u_batch_size=128)
        ds = BasicDataset(n, m, tu.numpy(), ti.numpy(), su.numpy(), si.numpy(), config=cfg, device=dev)
        gen = torch.Generator().manual_seed(3)
        E0 = (torch.randn(n + m, 64, generator=gen) * 0.1).to(dev)
        u = torch.randint(0, n, (B,), generator=gen).to(dev)
        p = torch.randint(0, m, (B,), generator=gen).to(dev)
        q = torch.randint(0, m, (B,), generator=gen).to(dev)
        sm = LightGCN(cfg, ds)
        with torch.no_grad():
            sm.all_embedding.weight.copy_(E0)
        sm.train()
        dm = DistLightGCN(cfg, ds, rank, world)
        dm.load_global_embedding(E0)
        light = dm.part.unshard(_gather(dm, dm.computer_local()))
        with torch.no_grad():
            ref = torch.cat(sm.computer())
        e_prop = float((light - ref).abs().max() / ref.abs().max())
        losses = [float(dm.fused_step(u, p, q)) for _ in range(2)]
        refl = [float(sm.stageOne(u, p, q)) for _ in range(2)]
        e_emb = float((dm.gather_embedding() - sm.all_embedding.weight.detach()).abs().max())
        ret[rank] = (bool(dm.reduce_mode), e_prop, losses, refl, e_emb)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.timeout(120)
def test_auto_partition_keeps_user_rows_home_on_a_user_heavy_graph_2gpu():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker_user_heavy, args=(world, port, ret), nprocs=world, join=True)
        for rank in range(world):
            reduce_mode, e_prop, losses, refl, e_emb = ret[rank]
            assert reduce_mode
            assert e_prop < 1e-5, e_prop
            for a, b in zip(losses, refl):
                assert abs(a - b) < 1e-5 * abs(b), (losses, refl)
            assert e_emb < 1e-6, e_emb
