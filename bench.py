#!/usr/bin/env python
"""Benchmark of the LightGCN hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ours:       one "step" = one fused train step (K-layer propagation + BPR + backward
            + Adam) on one batch of B=2048 triples of the cfg-2 graph
            (BASELINE.json configs[1]); `value` = nnz(A_hat) * layers / t_step with
            inputs resident in HBM, `e2e` = the same through LightGCN.stageOne()
            with pinned HOST triples (H2D inside) and a D2H read of the loss.
            The eval half of the metric (full-rank top-20 users/s) rides in "eval".
reference:  the reference's CPU path (torch.sparse.mm xK + autograd + Adam, the
            oracle port of model/lgcn.py:127-133) on this box's host cores.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

METRIC = "LightGCN train edges*layers/sec (prop+BPR)"
UNIT = "edges*layers/s"
CFG2 = dict(n_users=30000, m_items=41000, n_interactions=1_250_000, seed=2020, d=64, layers=3, batch=2048,
            lr=1e-4, decay=1e-7)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--storage", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eval", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "push", "nccl"],
                    help="N>1: push = all-gather fused into the SpMM epilogue over NVLink peer memory")
    ap.add_argument("--partition", default="auto", choices=["auto", "side_split", "two_sided"],
                    help="N>1: side_split = users on the first N/2 ranks, items on the rest (a row is only sent to "
                         "the other side); two_sided = every rank owns 1/N of both sides")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "hbm", "cfg3", "cfg4", "cfg5"],
                    help="hbm: 2.4M x 0.6M x 60M-edge graph, d=128 (table >> L2) for the honest HBM roofline; "
                         "cfg3: BASELINE configs[2], 10M x 2M x 500M edges, d=128 (graph built on device); "
                         "cfg4: configs[3], LightGCNSSM 4-layer d=64, 256 negatives per positive on the same graph; "
                         "cfg5: configs[4], full-rank top-20 eval of --eval-users users x 2M items, user-sharded")
    ap.add_argument("--eval-users", type=int, default=1_000_000, help="cfg5: users scored (all ranks together)")
    return ap.parse_args()


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return float(j["hbm_gbs"]), float(j.get("bf16_tflops_sustained", j.get("bf16_tflops", 1424.4))), "measured"
    return 6650.0, 1400.0, "fallback"  # B200_PROFILING.md fallback figures


def spmm_layer_bytes(nnz: int, N: int, d: int, s: int) -> int:
    """col ids + gathered rows + write Z + read/write fp32 layer sum + rowptr."""
    return nnz * (4 + d * s) + N * d * s + 2 * N * d * 4 + (N + 1) * 8


# ------------------------------------------------------------------ reference arm / cpu baseline
def cpu_train_steps(ds_arrays, steps: int, warmup: int, threads: int):
    """The reference's CPU path (oracle port): unsplit torch.sparse.mm graph, autograd, torch Adam."""
    import numpy as np
    import torch
    from oracle import lgcn_oracle as orc
    torch.set_num_threads(threads)
    n, m, tu, ti = ds_arrays
    g = torch.Generator().manual_seed(2020)
    E = torch.randn(n + m, CFG2["d"], generator=g) * 0.1
    om = orc.OracleModel(n, m, tu, ti, E, CFG2["layers"], CFG2["lr"], CFG2["decay"])
    rng = np.random.default_rng(0)
    B = CFG2["batch"]
    times = []
    for s in range(warmup + steps):
        idx = rng.integers(0, len(tu), B)
        users = torch.from_numpy(tu[idx]); pos = torch.from_numpy(ti[idx])
        neg = torch.from_numpy(rng.integers(0, m, B))
        t0 = time.perf_counter()
        om.stage_one(users, pos, neg)
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def workload_name(tag: str, world: int, K: int, d: int, B: int, n: int, m: int, nnz: int) -> str:
    """config.workload of both arms (the driver compares them)."""
    return (f"{tag}{' x%d' % world if world > 1 and tag == 'cfg-2' else ''}: LightGCN {K}-layer d={d} BPR B={B} on a synthetic "
            f"five-core bipartite graph {n} users x {m} items, nnz(A_hat)={nnz}")


def run_reference(args, rank: int):
    if rank != 0:
        return
    import torch
    from furusato_recommend_b200.synthetic import bipartite
    world = max(1, int(args.gpus))   # the same graph our arm trains on at this N (cfg-2 x N, weak scaling)
    n, m, tu, ti, su, si = bipartite(CFG2["n_users"] * world, CFG2["m_items"] * world, CFG2["n_interactions"] * world,
                                     seed=CFG2["seed"])
    nnz = 2 * int(tu.numel())
    threads = os.cpu_count() or 1
    steps = max(1, min(args.steps, 10))  # ~1 s per CPU step: keep the arm within a few minutes
    warm = max(1, min(args.warmup, 2))
    t = cpu_train_steps((n, m, tu.numpy(), ti.numpy()), steps, warm, threads)
    val = nnz * CFG2["layers"] / t
    sample = f"{steps} stageOne steps ({warm} warm-up) of the cfg-2 graph, unsplit torch.sparse.mm, B={CFG2['batch']}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name("cfg-2", world, CFG2["layers"], CFG2["d"], CFG2["batch"], n, m, nnz),
                   "path": "reference CPU path (oracle port, torch.sparse) on the host cores of rank 0"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------ our arm
def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from furusato_recommend_b200 import LightGCN, UniformSample, Trainer
    from furusato_recommend_b200.dataloader import BasicDataset
    from furusato_recommend_b200.synthetic import bipartite
    from furusato_recommend_b200 import ops, metric as lmetric

    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))

    # weak scaling: the graph grows with the number of GPUs (world x cfg-2), rows are
    # partitioned over the ranks (nnz-balanced) and every layer all-gathers over NVLink
    W = dict(CFG2)
    if args.workload == "hbm":
        W.update(n_users=2_400_000, m_items=600_000, n_interactions=75_000_000, d=128)
    if args.workload == "cfg3":
        W.update(n_users=10_000_000, m_items=2_000_000, n_interactions=625_000_000, d=128)
    if args.workload == "cfg4":
        W.update(n_users=10_000_000, m_items=2_000_000, n_interactions=625_000_000, d=64, layers=4, neg_size=256)
    cfg = dict(recdim=W["d"], layer=W["layers"], lr=W["lr"], decay=W["decay"],
               bpr_batch_size=W["batch"], device=dev, test_u_batch_size=10000, storage_dtype=args.storage,
               dist_exchange=args.exchange, dist_partition=args.partition)
    if args.workload in ("hbm", "cfg3", "cfg4"):
        from furusato_recommend_b200.dataloader import DeviceDataset
        n, m, tu, ti, su, si = bipartite(W["n_users"], W["m_items"], W["n_interactions"], seed=W["seed"], device=dev)
        ds = DeviceDataset(n, m, tu, ti, su, si, config=cfg)
        args.no_eval = True
        args.no_cpu_baseline = True
    else:
        n, m, tu, ti, su, si = bipartite(W["n_users"] * world, W["m_items"] * world,
                                         W["n_interactions"] * world, seed=W["seed"])
        ds = BasicDataset(n, m, tu.numpy(), ti.numpy(), su.numpy(), si.numpy(), config=cfg, device=dev)
    torch.manual_seed(2020 + rank)
    K, d, B = W["layers"], W["d"], W["batch"]
    J = int(W.get("neg_size", 1))
    if J > 1:
        if world != 1:
            raise SystemExit("cfg4 is measured on one GPU")
        from furusato_recommend_b200 import LightGCNSSM
        cfg["neg_size"] = J
        model = LightGCNSSM(cfg, ds)   # reference arithmetic: the BPR softplus over J*B flat triples per step
        model.train()
        nnz = model.graph.nnz
        fused = model._fused_step
        launches_per_step = 2 * K + 2
        B = B * J                      # rows per step (model/lgcnssm.py:141)
    elif world == 1:
        model = LightGCN(cfg, ds)
        model.train()
        nnz = model.graph.nnz
        fused = model._fused_step
        launches_per_step = 2 * K + 2
    else:
        from furusato_recommend_b200.parallel import DistLightGCN
        model = DistLightGCN(cfg, ds, rank, world)
        nnz = ds.csr_graph().nnz
        fused = model.fused_step
        # SpMM x 2K, BPR, Adam tick; push exchange adds 2 scale_rows_push + 1 exchange_rows_push
        launches_per_step = 2 * K + 2 + (3 if model.exchange == "push" else 0)
    N = n + m

    S = UniformSample(ds, neg_ratio=J, seed=CFG2["seed"], epoch=0,
                      count=min(ds.trainDataSize, (B // J) * (512 if J == 1 else 4)))
    n_batches = len(S) // B
    users, pos, neg = (S[:, j].contiguous() for j in range(3))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step(i):
        b = (i % n_batches) * B
        fused(users[b:b + B], pos[b:b + B], neg[b:b + B])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)            # L2 flush between timed iterations (untimed)
        ev[i][0].record()
        step(args.warmup + i)
        ev[i][1].record()
    barrier()
    t_dev = sum(a.elapsed_time(b) for a, b in ev) / 1e3  # seconds for K steps

    # ---- per-launch SpMM time, live, with events around every propagate launch ----
    spmm_ms = []
    orig = ops.propagate_layer

    def timed_layer(*a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); orig(*a, **kw); e1.record()
        spmm_ms.append((e0, e1))
    import furusato_recommend_b200.model as _m
    _m.ops.propagate_layer = timed_layer
    had_graph = getattr(model, "use_cuda_graph", False)
    if had_graph:
        model.use_cuda_graph = False      # the per-launch events need the eager launches
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        step(i)
    torch.cuda.synchronize()
    _m.ops.propagate_layer = orig
    if had_graph:
        model.use_cuda_graph = True
    spmm_avg_s = sum(a.elapsed_time(b) for a, b in spmm_ms) / len(spmm_ms) / 1e3

    # ---- e2e: public API with pinned host triples, H2D + loss D2H inside the timed region ----
    S_host = S.cpu()
    hu, hp, hn = (S_host[:, j].contiguous().pin_memory() for j in range(3))

    def e2e_step(b):
        if world == 1:
            return model.stageOne(hu[b:b + B], hp[b:b + B], hn[b:b + B]).item()
        return model.fused_step(hu[b:b + B], hp[b:b + B], hn[b:b + B]).item()

    for i in range(3):
        e2e_step(0)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        e2e_step((i % n_batches) * B)
    e1.record()
    barrier()
    t_e2e = e0.elapsed_time(e1) / 1e3
    # the timed regions are a few tens of ms; keep the same step running ~1.5 s more so the
    # 100 ms nvidia-smi sampler sees the clocks / throttle reasons of THIS workload under load
    t_end = time.perf_counter() + 1.5
    i = 0
    while time.perf_counter() < t_end:
        step(i)
        i += 1
        if i % 64 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    clk = clocks.stop()

    # ---- eval half of the metric: full-rank top-20 users/s ----
    eval_info = None
    if not args.no_eval and world == 1:
        model.eval()
        tr = Trainer(cfg, ds, model)
        ev_users = tr._eval_users()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(); model.computer(); p1.record()
        rp, srt = ds.test_csr()
        peak_tf = measured_peaks()[1]

        def eval_pass(precision):
            sums = torch.zeros((4, 2), dtype=torch.float64, device=dev)
            idx, _ = model.getUsersTopK(ev_users, 20, precision=precision)   # one launch: nothing is materialised
            lmetric.batch_metric_sums(idx, ev_users, rp, srt, (10, 20), sums)
            return sums

        def timed(fn, reps):
            fn(); torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(reps):
                fn()
            a1.record(); torch.cuda.synchronize()
            return a0.elapsed_time(a1) / 1e3 / reps

        t_tc = timed(lambda: eval_pass("bf16"), 5)
        t_f32 = timed(lambda: eval_pass("fp32"), 3)
        model.eval_precision = "bf16"
        tr.test()
        w0 = time.perf_counter()
        res = tr.test()  # public API: host user ids in, metric dict out
        torch.cuda.synchronize()
        t_eval_e2e = time.perf_counter() - w0
        model.eval_precision = "fp32"
        res32 = tr.test()
        flops = 2.0 * len(ev_users) * m * d
        eval_info = {"metric": "full-rank top-20 eval users/sec (score+mask+top-k+metrics)",
                     "value": len(ev_users) / t_tc, "unit": "users/s", "users": int(len(ev_users)), "items": m,
                     "precision": "bf16 tcgen05 (fp32 accumulate)", "propagation_ms": p0.elapsed_time(p1),
                     "e2e": {"value": len(ev_users) / t_eval_e2e, "unit": "users/s"},
                     "tflops": flops / t_tc / 1e12, "tensor_peak_tflops": peak_tf,
                     "fp32_exact_value": len(ev_users) / t_f32,
                     "recall@20": float(res["recall"][1]), "ndcg@20": float(res["ndcg"][1]),
                     "recall@20_fp32": float(res32["recall"][1]), "ndcg@20_fp32": float(res32["ndcg"][1])}
        # cfg-5-shaped sweep: the tensor-pipe figure is only meaningful at m = 2M items
        if not args.no_sweep:
            U5, m5 = 148 * 128 * 4, 2_000_000
            g5 = torch.Generator(device=dev).manual_seed(5)
            ue5 = torch.randn(U5, d, generator=g5, device=dev) * 0.1
            ie5 = torch.randn(m5, d, generator=g5, device=dev) * 0.1
            ids5 = torch.arange(U5, device=dev)
            npos = 50
            rp5 = torch.arange(U5 + 1, device=dev, dtype=torch.int64) * npos
            pos5 = torch.sort(torch.randint(0, m5, (U5, npos), generator=g5, device=dev, dtype=torch.int32), dim=1)[0]
            pos5 = pos5.reshape(-1).contiguous()
            t5 = timed(lambda: ops.score_topk(ue5, ie5, ids5, rp5, pos5, 20, precision="bf16"), 3)
            fl5 = 2.0 * U5 * m5 * d
            eval_info["sweep_cfg5"] = {"users": U5, "items": m5, "d": d, "k": 20, "masked_per_user": npos,
                                       "value": U5 / t5, "unit": "users/s", "ms": t5 * 1e3,
                                       "tflops": fl5 / t5 / 1e12, "frac_of_tensor_peak": fl5 / t5 / 1e12 / peak_tf,
                                       "includes": "fp32->bf16 operand packing of both tables + score + mask + top-k"}
            del ue5, ie5, pos5
        model.train()

    # ---- max over ranks ----
    if world > 1:
        t = torch.tensor([t_dev, t_e2e, spmm_avg_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, t_e2e, spmm_avg_s = (float(x) for x in t)
    nnz_total = float(nnz)  # the partitioned graph is ONE graph: every rank reports its global nnz
    dist_eval = None
    if world > 1 and not args.no_eval:
        ev_users = torch.from_numpy(ds.test_users()).to(dev)
        model.computer_local()
        model.topk_user_shard(ev_users, 20)
        barrier()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        mine, idx, _ = model.topk_user_shard(ev_users, 20)
        q1.record()
        barrier()
        te = torch.tensor([q0.elapsed_time(q1) / 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist_eval = {"metric": "full-rank top-20 eval users/sec, user-sharded (incl. all-gather of light_out)",
                     "value": len(ev_users) / float(te), "unit": "users/s", "users": int(len(ev_users)), "items": m,
                     "precision": "bf16 tcgen05 (fp32 accumulate)"}

    if rank == 0:
        hbm_peak, _, which = measured_peaks()
        s_bytes = 2 if args.storage == "bf16" else 4
        layer_bytes = spmm_layer_bytes(nnz // world, N // world, d, s_bytes)  # per rank, per launch
        achieved = layer_bytes / spmm_avg_s / 1e9
        n_loc, nnz_loc = N // world, nnz // world
        compulsory = nnz_loc * 4 + (n_loc + 1) * 8 + 2 * n_loc * d * s_bytes
        gather_ceiling = 19800.0 if n_loc * d * s_bytes < 100e6 else 7200.0
        traffic = None
        tp = REPO / "profiles" / "spmm_traffic.json"
        if tp.exists():
            traffic = json.loads(tp.read_text()).get(f"{args.workload}_dram_bytes_per_launch_{args.storage}")
        ms = t_dev / args.steps * 1e3
        out = {
            "metric": METRIC, "value": nnz_total * K / (t_dev / args.steps), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak" if args.workload == "cfg2" else "strong", "vs_baseline": None,
            "dtype": "f32" if args.storage == "fp32" else "bf16-storage/f32-acc",
            "data": "synthetic",
            "config": {"workload": workload_name({'hbm': 'hbm-bound', 'cfg3': 'cfg-3', 'cfg2': 'cfg-2',
                                                  'cfg4': 'cfg-4 (lgcnssm, %d negatives per positive)' % J}[args.workload],
                                                 world, K, d, B, n, m, nnz),
                       "l2": "256 MiB buffer written between timed steps (L2 flush, untimed)",
                       "parallelism": "1 GPU" if world == 1 else
                       f"{world} GPUs: rows partitioned by nnz; per-layer exchange = " +
                       ("all-gather fused into the SpMM epilogue (NVLink peer stores, symmetric memory; "
                        + ("bipartite side split: a row goes to the W/2 ranks of the other side only"
                           if getattr(model.part, "side_split", False) else "every row to every rank")
                        + ") (2K per step) + 3B-row exchange by owner peer stores; the whole step is one CUDA graph"
                        if getattr(model, "exchange", "") == "push" else
                        "ncclAllGather (2K per step) + one 3B-row all-reduce")},
            "e2e": {"value": nnz_total * K / (t_e2e / args.steps), "unit": UNIT, "h2d_bytes_per_step": 3 * B * 8,
                    "d2h_bytes_per_step": 4, "ms_per_step": t_e2e / args.steps * 1e3},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "hbm", "kernel": "spmm_layer_kernel", "achieved": achieved, "peak": hbm_peak,
                         "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": which,
                         "bytes_per_launch": layer_bytes, "avg_launch_us": spmm_avg_s * 1e6,
                         "launches_per_step": 2 * K,
                         # SURVEY 8d caveat: with an L2-resident table the gather term never reaches DRAM, so also
                         # report the compulsory DRAM bytes (col ids + rowptr + read/write of the [N, d] rows) and the
                         # measured ceiling of this access pattern (tools/gather_ceiling.cu: 19.8 TB/s from L2,
                         # 7.2 TB/s from HBM) the kernel is really up against
                         "compulsory_bytes_per_launch": compulsory,
                         "compulsory_gbs": compulsory / spmm_avg_s / 1e9,
                         "gather_ceiling_gbs": gather_ceiling,
                         "frac_of_gather_ceiling": achieved / gather_ceiling,
                         "note": ("gather model; the table is %.0f MB" % (N * d * s_bytes / 1e6)) +
                                 (" (L2-resident, so frac may exceed 1)" if N * d * s_bytes < 100e6 else " (>> 126 MB L2)")},
            "clocks": clk,
        }
        if eval_info:
            out["eval"] = eval_info
        if dist_eval:
            out["eval"] = dist_eval
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            t_cpu = cpu_train_steps((n, m, tu.cpu().numpy(), ti.cpu().numpy()), 5, 2, threads)
            out["cpu_baseline"] = {"value": nnz * K / t_cpu, "unit": UNIT, "cores": threads, "kind": "port",
                                   "sample": "5 stageOne steps (2 warm-up) of the same cfg-2 graph on the host: unsplit "
                                             "torch.sparse.mm x3 + autograd + torch Adam, B=2048",
                                   "ms_per_step": t_cpu * 1e3}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------ cfg-5: eval sweep
def run_cfg5(args, rank: int, world: int, local_rank: int):
    """BASELINE configs[4]: full-rank top-20 over 2M items, d=64, ~50 masked train positives per
    user, users sharded over the ranks (independent units: no data-path collective; the item table
    is replicated).  One step = score + mask + top-k of this rank's user shard (operand packing
    included); value = users of all ranks / max-over-ranks time."""
    import torch
    import torch.distributed as dist
    from furusato_recommend_b200 import ops

    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    d, k, npos, m5 = 64, 20, 50, 2_000_000
    U = args.eval_users // world
    g5 = torch.Generator(device=dev).manual_seed(5 + rank)
    gi = torch.Generator(device=dev).manual_seed(5)
    ie = torch.randn(m5, d, generator=gi, device=dev) * 0.1           # same item table on every rank
    ue = torch.randn(U, d, generator=g5, device=dev) * 0.1
    ids = torch.arange(U, device=dev)
    rp = torch.arange(U + 1, device=dev, dtype=torch.int64) * npos
    pos = torch.empty((U, npos), dtype=torch.int32, device=dev)
    for a in range(0, U, 1 << 20):                                     # bounded temporaries
        b = min(U, a + (1 << 20))
        pos[a:b] = torch.sort(torch.randint(0, m5, (b - a, npos), generator=g5, device=dev, dtype=torch.int32), dim=1)[0]
    pos = pos.reshape(-1)
    chunk = 148 * 128 * 8                                              # users per launch (8 CTA waves)

    def step():
        out = []
        for a in range(0, U, chunk):
            out.append(ops.score_topk(ue, ie, ids[a:a + chunk], rp, pos, k, precision="bf16")[0])
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    steps, warm = max(1, min(args.steps, 3)), max(1, min(args.warmup, 3))
    for _ in range(warm):
        step()
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    t = e0.elapsed_time(e1) / 1e3 / steps
    # e2e: host user ids in (pinned), top-k ids back on the host
    hid = ids.cpu().pin_memory()
    hout = torch.empty((U, k), dtype=torch.int32).pin_memory()
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0.record()
    for a in range(0, U, chunk):
        did = hid[a:a + chunk].to(dev, non_blocking=True)
        hout[a:a + chunk].copy_(ops.score_topk(ue, ie, did, rp, pos, k, precision="bf16")[0], non_blocking=True)
    w1.record()
    barrier()
    t_e2e = w0.elapsed_time(w1) / 1e3
    clk = clocks.stop()
    if world > 1:
        tt = torch.tensor([t, t_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t, t_e2e = float(tt[0]), float(tt[1])
    if rank == 0:
        _, peak_tf, which = measured_peaks()
        flops = 2.0 * U * m5 * d
        n_launch = (U + chunk - 1) // chunk
        print(json.dumps({
            "metric": "full-rank top-20 eval users/sec", "value": U * world / t, "unit": "users/s", "n_gpus": world,
            "steps": steps, "warmup": warm, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16 operands / f32 accumulate (tcgen05)", "data": "synthetic",
            "config": {"workload": f"cfg-5: full-rank top-{k} eval, {U * world} users x {m5} items, d={d}, {npos} masked "
                                   f"train positives per user, users sharded over {world} GPU(s)",
                       "l2": "operands (>= 256 MB item table) exceed L2; no flush needed",
                       "parallelism": f"{world} GPU(s): user shards, item table replicated, no collective"},
            "e2e": {"value": U * world / t_e2e, "unit": "users/s", "h2d_bytes_per_step": U * 8, "d2h_bytes_per_step": U * k * 4},
            "gpu_launches": 3 * n_launch * steps,
            "roofline": {"bound": "tensor", "kernel": "score_topk_tc_kernel", "achieved": flops / t / 1e12, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": flops / t / 1e12 / peak_tf, "traffic": None, "peak_source": which,
                         "note": "2*U*m*d flops per rank; time includes the fp32->bf16 operand packing kernels"},
            "clocks": clk}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    if args.workload == "cfg5":
        run_cfg5(args, rank, world, local_rank)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
