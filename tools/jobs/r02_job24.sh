#!/bin/bash
# round-2 job 24 (1 GPU): candidate path v3 (argmax descent, unsorted top-k set with min scan, positives prefetch) — parity, A-B, phase profile
O=gpurun_out/r02x; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_cfg2.py tests/test_gpu_parity.py -q -x > $O/test_tc.log 2>&1; echo "rc=$?" >> $O/test_tc.log
LGCN_TC_LAYOUT=m2rl timeout 600 python -m pytest tests/test_gpu_tc.py -q -x > $O/test_tc_m2rl.log 2>&1; echo "rc=$?" >> $O/test_tc_m2rl.log
SW="timeout 120 python tools/topk_sweep.py --users 75776 --items 2000000"
P=$PWD/furusato_recommend_b200/liblgcn_b200_tcprof.so
$SW > $O/sweep_default.log 2>&1
$SW --k 1 > $O/sweep_k1.log 2>&1
$SW --k 10 > $O/sweep_k10.log 2>&1
$SW --d 128 > $O/sweep_d128.log 2>&1
$SW --npos 0 > $O/sweep_npos0.log 2>&1
$SW --npos 200 > $O/sweep_npos200.log 2>&1
LGCN_TC_TRIG=23 $SW > $O/sweep_trig23.log 2>&1
LGCN_TC_TRIG=30 $SW > $O/sweep_trig30.log 2>&1
LGCN_TC_TRIG=36 $SW > $O/sweep_trig36.log 2>&1
LGCN_TC_LAYOUT=m2rl $SW > $O/sweep_m2rl.log 2>&1
LGCN_B200_LIB=$P $SW > $O/prof_default.log 2>&1
tail -n 3 $O/test_tc.log $O/test_tc_m2rl.log; for f in $O/sweep_*.log; do echo "$(basename $f .log): $(tail -n 1 $f)"; done; head -n 5 $O/prof_default.log | cut -c1-300
