"""Where does the A_hat sqrt(deg) identity deviate under a given partition?

  torchrun --nproc-per-node W tools/debug_reduce_parity.py [--shape hbm|cfg3] [--partition reduce|two_sided|auto]
                                                           [--own-graph]

With E = sqrt(deg) (x) v every layer of the propagation must return E itself, so any row of computer() that differs
from the table localises an error.  Prints, per rank and side, the worst row (error relative to the largest table
entry, its degree) and the number of rows whose error exceeds 1e-5 of their own magnitude — for a fresh model, after
six train steps, and once more.  `--own-graph` makes every rank generate its own copy of the synthetic graph (what
bench.py did before `shared_device_graph`): that is how the cross-rank graph mismatch showed up as "one missing
neighbour" on a few item rows under the reduce partition (profiles/r02_reduce_parity_debug.log).
"""
import argparse
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from furusato_recommend_b200.dataloader import DeviceDataset  # noqa: E402
from furusato_recommend_b200.parallel import DistLightGCN  # noqa: E402
from furusato_recommend_b200.synthetic import bipartite  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="hbm", choices=["hbm", "cfg3"])
ap.add_argument("--partition", default="reduce")
ap.add_argument("--own-graph", action="store_true")
a = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = f"cuda:{rank}"
dist.init_process_group("nccl", device_id=torch.device(dev))
W = dict(bench.CFG3 if a.shape == "cfg3" else bench.HBM)
if a.own_graph:
    n, m, tu, ti, su, si = bipartite(W["n_users"], W["m_items"], W["n_interactions"], seed=W["seed"], device=dev)
else:
    n, m, tu, ti, su, si = bench.shared_device_graph(W, dev, rank, world)


def report(model, deg, K, phase):
    v = torch.linspace(0.5, 1.5, W["d"], device=dev)
    model.load_global_embedding(deg.sqrt()[:, None] * v[None, :])
    out = model.computer_local()
    err = (out - model.emb).abs().amax(dim=1)             # per local row
    scale = float(deg.max().sqrt() * 1.5)
    (u_lo, u_hi), (i_lo, i_hi) = model.part.ranges(rank)
    nu, ni = u_hi - u_lo, i_hi - i_lo
    eu, ei = err[:nu], err[nu:nu + ni]
    ju, ji = int(eu.argmax()), int(ei.argmax())
    bad_u = eu > 1e-5 * model.emb[:nu].abs().amax(dim=1)
    bad_i = ei > 1e-5 * model.emb[nu:nu + ni].abs().amax(dim=1)
    msg = (f"K={K} {phase} rank {rank}/{world} {a.partition}: users max {float(eu[ju]) / scale:.2e} of the largest entry "
           f"(deg {int(deg[u_lo + ju])}, local row {ju} of {nu}) | items max {float(ei[ji]) / scale:.2e} "
           f"(deg {int(deg[i_lo + ji])}, local row {ji} of {ni}); row-relative > 1e-5: users {int(bad_u.sum())} "
           f"items {int(bad_i.sum())}")
    for r in range(world):
        if r == rank:
            print(msg, flush=True)
        dist.barrier()


for K in (1, 3):
    cfg = dict(recdim=W["d"], layer=K, lr=1e-4, decay=1e-7, bpr_batch_size=2048, device=dev, dist_partition=a.partition)
    ds = DeviceDataset(n, m, tu, ti, su, si, config=cfg)
    g = ds.csr_graph()
    deg = (g.rowptr[1:] - g.rowptr[:-1]).to(torch.float32)
    model = DistLightGCN(cfg, ds, rank, world)
    report(model, deg, K, "fresh")
    gen = torch.Generator(device=dev).manual_seed(1)
    for _ in range(6):
        u = torch.randint(0, n, (2048,), generator=gen, device=dev)
        p = torch.randint(0, m, (2048,), generator=gen, device=dev)
        q = torch.randint(0, m, (2048,), generator=gen, device=dev)
        model.fused_step(u, p, q)
    report(model, deg, K, "after 6 train steps")
    report(model, deg, K, "again")
    del model, ds
    torch.cuda.empty_cache()
dist.destroy_process_group()
