"""Where does the set-up time of the big synthetic workloads go?  (generator, CSR build, work lists)"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from furusato_recommend_b200.synthetic import bipartite  # noqa: E402
from furusato_recommend_b200.graph import build_csr_graph, build_pos_csr  # noqa: E402

n_u, m_i, n_int = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (10_000_000, 2_000_000, 625_000_000)))
dev = "cuda:0"
torch.cuda.init()


def tick(label, t0):
    torch.cuda.synchronize()
    t = time.perf_counter()
    print(f"{label}: {t - t0:.2f} s  (peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB)", flush=True)
    return t


t = time.perf_counter()
n, m, tu, ti, su, si = bipartite(n_u, m_i, n_int, seed=2020, device=dev)
t = tick(f"bipartite -> n={n} m={m} train={tu.numel()} test={su.numel()}", t)
g = build_csr_graph(n, m, tu, ti)
t = tick(f"build_csr_graph nnz={g.nnz} n_light={g.light_rows.numel()} n_seg={g.seg_row.numel()}", t)
p = build_pos_csr(n, tu, ti)
t = tick("build_pos_csr", t)
