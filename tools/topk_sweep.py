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
# tuning builds (-DLGCN_TC_PROF=1, selected with LGCN_B200_LIB): per-phase clock totals of CTA 0
import ctypes  # noqa: E402
from furusato_recommend_b200 import _lib  # noqa: E402
try:
    fn = _lib.load().lgcn_debug_tc_prof
except AttributeError:
    fn = None
if fn is not None:
    buf = (ctypes.c_longlong * 24)()
    fn(buf)
    v = list(buf)
    nt = max(v[5], 1)
    print("epilogue warp, clocks per item tile: wait-full %.0f  tmem-read %.0f  tree+vote %.0f  candidates %.0f  hand-back %.0f  "
          "total %.0f  (entries %d = %.3f per tile, %.0f clk per entry)" %
          (v[0] / nt, v[1] / nt, v[2] / nt, v[3] / nt, v[4] / nt, v[7] / nt, v[6], v[6] / nt, v[3] / max(v[6], 1)))
    ne = max(v[6], 1)
    print("candidate path per entry: hit masks %.0f  single take %.0f  multi path %.0f  trigger vote + compaction %.0f   "
          "(multi entries %d, compactions %d)" % (v[16] / ne, v[17] / ne, v[18] / ne, v[19] / ne, v[20], v[21]))
    nt = max(v[11], 1)
    print("issuer warp, clocks per item tile: wait-B-tile %.0f  wait-acc-empty %.0f  issue %.0f  total %.0f" %
          (v[8] / nt, v[9] / nt, v[10] / nt, v[12] / nt))
    try:
        fs = _lib.load().lgcn_debug_tc_stamps
    except AttributeError:
        fs = None
    if fs is not None:
        sb = (ctypes.c_longlong * 40)()
        fs(sb)
        s = [list(sb[5 * i:5 * i + 5]) for i in range(8)]
        t0 = min(x for row in s for x in row if x > 0)
        print("tile: acc-free  issued | acc-full  read-done  handed-back   (clocks since the first stamp; issuer 0 / epilogue warp 0 of CTA 0)")
        for i, row in enumerate(s):
            print("  %d: %6d %6d | %6d %6d %6d" % tuple([i] + [x - t0 for x in row]))
