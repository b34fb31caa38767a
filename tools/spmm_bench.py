"""Per-launch timing of the SpMM layer kernel and the whole fused step at cfg-2 (tuning aid).
Select a library variant with LGCN_B200_LIB=<path to .so>."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from furusato_recommend_b200 import LightGCN, UniformSample, ops  # noqa: E402
from furusato_recommend_b200.dataloader import BasicDataset  # noqa: E402
from furusato_recommend_b200.synthetic import bipartite  # noqa: E402

d = int(os.environ.get("D", 64))
storage = os.environ.get("STORAGE", "fp32")
n, m, tu, ti, su, si = bipartite(30000, 41000, 1_250_000, seed=2020)
cfg = dict(recdim=d, layer=3, lr=1e-4, decay=1e-7, bpr_batch_size=2048, device="cuda:0", storage_dtype=storage,
           cuda_graph=True)
ds = BasicDataset(n, m, tu.numpy(), ti.numpy(), su.numpy(), si.numpy(), config=cfg, device="cuda:0")
model = LightGCN(cfg, ds)
model.train()
S = UniformSample(ds, seed=1, epoch=0)
u, p, q = (S[:2048, j].contiguous() for j in range(3))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
g = model.graph
E = model.all_embedding.weight.data
z0, z1, acc = model._buf("Z0"), model._buf("Z1"), model._buf("ACC")


def timeit(fn, reps=30, do_flush=True):
    for _ in range(5):
        fn()
    tot = 0.0
    for _ in range(reps):
        if do_flush:
            flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps * 1e3


first = timeit(lambda: ops.propagate_layer(g, E, scale_src=True, dst=z0, acc_in=E, acc_out=acc))
mid = timeit(lambda: ops.propagate_layer(g, z0, scale_src=False, dst=z1, acc_in=acc, acc_out=acc))
mid_warm = timeit(lambda: ops.propagate_layer(g, z0, scale_src=False, dst=z1, acc_in=acc, acc_out=acc), do_flush=False)
step = timeit(lambda: model._fused_step(u, p, q))
step_warm = timeit(lambda: model._fused_step(u, p, q), do_flush=False)
print(f"{os.environ.get('LGCN_B200_LIB', 'default'):>40s} d={d} {storage}: first {first:6.1f} us  mid {mid:6.1f} us "
      f"(warm {mid_warm:6.1f})  step {step:6.1f} us (warm {step_warm:6.1f})")
