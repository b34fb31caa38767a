#!/bin/bash
# round-2 job 20 (1 GPU): register-staged + lane-local top-k epilogue (layout m2rl) — parity, then the debug ladder A-B against m2g2
O=gpurun_out/r02t; mkdir -p $O
LGCN_TC_LAYOUT=m2rl timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_cfg2.py -q -x > $O/test_tc_m2rl.log 2>&1; echo "rc=$?" >> $O/test_tc_m2rl.log
SW="timeout 120 python tools/topk_sweep.py --users 75776 --items 2000000"
$SW > $O/sweep_m2g2.log 2>&1
LGCN_TC_LAYOUT=m2rl $SW > $O/sweep_m2rl.log 2>&1
LGCN_TC_LAYOUT=m2rl LGCN_TC_DEBUG=1 $SW > $O/sweep_m2rl_dbg1.log 2>&1
LGCN_TC_LAYOUT=m2rl LGCN_TC_DEBUG=2 $SW > $O/sweep_m2rl_dbg2.log 2>&1
LGCN_TC_LAYOUT=m2rl LGCN_TC_DEBUG=3 $SW > $O/sweep_m2rl_dbg3.log 2>&1
LGCN_TC_LAYOUT=m2rl $SW --k 1 > $O/sweep_m2rl_k1.log 2>&1
LGCN_TC_LAYOUT=m2rl $SW --k 10 > $O/sweep_m2rl_k10.log 2>&1
LGCN_TC_LAYOUT=m2rl LGCN_TC_TRIG=24 $SW > $O/sweep_m2rl_trig24.log 2>&1
LGCN_TC_LAYOUT=m2rl LGCN_TC_TRIG=30 $SW > $O/sweep_m2rl_trig30.log 2>&1
LGCN_TC_LAYOUT=m2rl LGCN_TC_TRIG=36 $SW > $O/sweep_m2rl_trig36.log 2>&1
LGCN_TC_LAYOUT=m2rl LGCN_TC_TRIG=44 $SW > $O/sweep_m2rl_trig44.log 2>&1
tail -n 3 $O/test_tc_m2rl.log; tail -qn 1 $O/sweep_*.log
# SpMM on the HBM-bound graph (2.4M x 0.6M x 60M edges, d=128): UNROLL 8 variants against the shipped UNROLL 4
for st in bf16 fp32; do
for lib in liblgcn_b200.so liblgcn_b200_u8.so liblgcn_b200_u8b6.so; do
LGCN_B200_LIB=$PWD/furusato_recommend_b200/$lib timeout 300 python bench.py --workload hbm --storage $st --steps 10 --warmup 3 > $O/hbm_${st}_${lib%.so}.log 2> $O/hbm_${st}_${lib%.so}.err
python - <<PY
import json
try:
    j = json.loads(open("$O/hbm_${st}_${lib%.so}.log").read().strip().splitlines()[-1])
    print("$st $lib", "ms_per_step", j["ms_per_step"], "spmm_us", j.get("roofline", {}).get("avg_launch_us"), j.get("roofline", {}).get("frac"))
except Exception as e:
    print("$st $lib failed", e)
PY
done
done
