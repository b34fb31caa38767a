#!/bin/bash
# round-2 job 27 (1 GPU): 16 epilogue warps (m2g4: two column groups per user tile) with the v3 candidate path
O=gpurun_out/r02aa; mkdir -p $O
SW="timeout 120 python tools/topk_sweep.py --users 75776 --items 2000000"
$SW > $O/sweep_m2g2.log 2>&1
LGCN_TC_LAYOUT=m2g4 $SW > $O/sweep_m2g4.log 2>&1
LGCN_TC_LAYOUT=m2g4 LGCN_TC_DEBUG=3 $SW > $O/sweep_m2g4_dbg3.log 2>&1
LGCN_TC_LAYOUT=m2g4 LGCN_TC_DEBUG=1 $SW > $O/sweep_m2g4_dbg1.log 2>&1
LGCN_TC_LAYOUT=m2g4 $SW --k 10 > $O/sweep_m2g4_k10.log 2>&1
LGCN_TC_LAYOUT=m2g4 $SW --k 1 > $O/sweep_m2g4_k1.log 2>&1
LGCN_TC_LAYOUT=m2g4 timeout 600 python -m pytest tests/test_gpu_tc.py -q -x > $O/test_tc_m2g4.log 2>&1; echo "rc=$?" >> $O/test_tc_m2g4.log
for f in $O/sweep_*.log; do echo "$(basename $f .log): $(tail -n 1 $f)"; done; tail -n 3 $O/test_tc_m2g4.log
