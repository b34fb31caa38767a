"""Is synthetic.bipartite() bit-reproducible across ranks and across calls?  torchrun --nproc-per-node W this.py"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from furusato_recommend_b200.synthetic import bipartite  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = f"cuda:{rank}"
dist.init_process_group("nccl", device_id=torch.device(dev))
W = dict(bench.CFG3 if (len(sys.argv) < 2 or sys.argv[1] == "cfg3") else bench.HBM)


def digest():
    n, m, tu, ti, su, si = bipartite(W["n_users"], W["m_items"], W["n_interactions"], seed=W["seed"], device=dev)
    h = torch.stack([torch.tensor(n, device=dev), torch.tensor(m, device=dev), torch.tensor(tu.numel(), device=dev),
                     (tu * 1000003 + ti * 7919).sum(), (tu * ti % 1000000007).sum(), (su * 31 + si).sum()]).to(torch.int64)
    del tu, ti, su, si
    torch.cuda.empty_cache()
    return h


a = digest()
b = digest()
allh = [torch.zeros_like(a) for _ in range(world)]
dist.all_gather(allh, a)
if rank == 0:
    print("same on every rank:", all(bool((x == allh[0]).all()) for x in allh))
    for r, x in enumerate(allh):
        print(r, x.tolist())
print(f"rank {rank}: two calls agree: {bool((a == b).all())}", flush=True)
dist.destroy_process_group()
