"""Where does the A_hat sqrt(deg) identity deviate under a given partition?  torchrun --nproc-per-node W this.py
[--shape hbm|cfg3] [--partition reduce|two_sided].  Prints, per side, the max error of every layer count 1..K and
the degree of the worst row."""
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
ap.add_argument("--shape", default="hbm")
ap.add_argument("--partition", default="reduce")
a = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = f"cuda:{rank}"
dist.init_process_group("nccl", device_id=torch.device(dev))
W = dict(bench.CFG3 if a.shape == "cfg3" else bench.HBM)
n, m, tu, ti, su, si = bipartite(W["n_users"], W["m_items"], W["n_interactions"], seed=W["seed"], device=dev)
for K in (1, 3):
    cfg = dict(recdim=W["d"], layer=K, lr=1e-4, decay=1e-7, bpr_batch_size=2048, device=dev, dist_partition=a.partition)
    ds = DeviceDataset(n, m, tu, ti, su, si, config=cfg)
    g = ds.csr_graph()
    deg = (g.rowptr[1:] - g.rowptr[:-1]).to(torch.float32)
    model = DistLightGCN(cfg, ds, rank, world)
    v = torch.linspace(0.5, 1.5, W["d"], device=dev)
    for phase in ("fresh", "after 6 train steps", "again"):
      if phase == "after 6 train steps":
        gen = torch.Generator(device=dev).manual_seed(1)
        for _ in range(6):
            u = torch.randint(0, n, (2048,), generator=gen, device=dev)
            p_ = torch.randint(0, m, (2048,), generator=gen, device=dev)
            q_ = torch.randint(0, m, (2048,), generator=gen, device=dev)
            model.fused_step(u, p_, q_)
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
      msg = (f"K={K} {phase} rank {rank}/{world} {a.partition}: users max {float(eu[ju]) / scale:.2e} of global max (deg {int(deg[u_lo + ju])}, "
             f"local row {ju} of {nu}) | items max {float(ei[ji]) / scale:.2e} (deg {int(deg[i_lo + ji])}, local row {ji} of {ni}); "
             f"row-relative > 1e-5: users {int(bad_u.sum())} items {int(bad_i.sum())}; pad rows max |out| {float(out[nu + ni:].abs().max()) if out.shape[0] > nu + ni else 0:.2e}")
      for r in range(world):
        if r == rank:
            print(msg, flush=True)
        dist.barrier()
    del model, ds
    torch.cuda.empty_cache()
    continue
    out = model.computer_local()
    err = (out - model.emb).abs().amax(dim=1)             # per local row
    scale = float(deg.max().sqrt() * 1.5)
    (u_lo, u_hi), (i_lo, i_hi) = model.part.ranges(rank)
    nu, ni = u_hi - u_lo, i_hi - i_lo
    eu, ei = err[:nu], err[nu:nu + ni]
    ju, ji = int(eu.argmax()), int(ei.argmax())
    rel_row_u = float(eu[ju] / model.emb[ju].abs().max().clamp_min(1e-30))
    rel_row_i = float(ei[ji] / model.emb[nu + ji].abs().max().clamp_min(1e-30))
    msg = (f"K={K} rank {rank}/{world} {a.partition}: users max abs err {float(eu[ju]):.3e} (/global max {float(eu[ju]) / scale:.2e}, "
           f"row-relative {rel_row_u:.2e}, deg {int(deg[u_lo + ju])}) | items max abs err {float(ei[ji]):.3e} "
           f"(/global max {float(ei[ji]) / scale:.2e}, row-relative {rel_row_i:.2e}, deg {int(deg[i_lo + ji])}); "
           f"rows with row-relative err > 1e-5: users {int((eu > 1e-5 * model.emb[:nu].abs().amax(dim=1)).sum())} items "
           f"{int((ei > 1e-5 * model.emb[nu:nu + ni].abs().amax(dim=1)).sum())}")
    for r in range(world):
        if r == rank:
            print(msg, flush=True)
        dist.barrier()
    del model, ds
    torch.cuda.empty_cache()
dist.destroy_process_group()
