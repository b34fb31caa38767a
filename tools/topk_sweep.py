"""Stand-alone timing of the fused score+mask+top-k kernels on a cfg-5-shaped problem
(random N(0,0.1) embeddings, `npos` masked train positives per user).  Used for ncu captures."""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from furusato_recommend_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--users", type=int, default=148 * 128)
ap.add_argument("--items", type=int, default=500_000)
ap.add_argument("--d", type=int, default=64)
ap.add_argument("--k", type=int, default=20)
ap.add_argument("--npos", type=int, default=50)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(5)
ue = torch.randn(a.users, a.d, generator=g, device=dev) * 0.1
ie = torch.randn(a.items, a.d, generator=g, device=dev) * 0.1
ids = torch.arange(a.users, device=dev)
rp = torch.arange(a.users + 1, device=dev, dtype=torch.int64) * a.npos
if a.npos > 0:
    pos = torch.sort(torch.randint(0, a.items, (a.users, a.npos), generator=g, device=dev, dtype=torch.int32), dim=1)[0].reshape(-1).contiguous()
else:   # no masked positives at all (isolates the cost of the positives walk)
    pos = torch.zeros(1, dtype=torch.int32, device=dev)
ops.score_topk(ue, ie, ids, rp, pos, a.k, precision=a.precision)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    ops.score_topk(ue, ie, ids, rp, pos, a.k, precision=a.precision)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
fl = 2.0 * a.users * a.items * a.d
print(f"users={a.users} items={a.items} d={a.d} k={a.k} {a.precision}: {ms:.3f} ms  {a.users / ms * 1e3:.3e} users/s  {fl / ms / 1e9:.1f} TFLOP/s")
