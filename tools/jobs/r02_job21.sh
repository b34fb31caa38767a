#!/bin/bash
# round-2 job 21 (1 GPU): single-hit shortcut in the top-k candidate path (m2g2) — parity, A-B, and per-phase clock
# totals of one epilogue / one issuing warp from the -DLGCN_TC_PROF build
O=gpurun_out/r02u; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_cfg2.py -q -x > $O/test_tc.log 2>&1; echo "rc=$?" >> $O/test_tc.log
SW="timeout 120 python tools/topk_sweep.py --users 75776 --items 2000000"
P=$PWD/furusato_recommend_b200/liblgcn_b200_tcprof.so
$SW > $O/sweep_default.log 2>&1
LGCN_TC_DEBUG=3 $SW > $O/sweep_default_dbg3.log 2>&1
$SW --k 1 > $O/sweep_k1.log 2>&1
$SW --k 10 > $O/sweep_k10.log 2>&1
$SW --d 128 > $O/sweep_d128.log 2>&1
LGCN_B200_LIB=$P $SW > $O/prof_default.log 2>&1
LGCN_B200_LIB=$P LGCN_TC_DEBUG=3 $SW > $O/prof_dbg3.log 2>&1
LGCN_B200_LIB=$P LGCN_TC_DEBUG=1 $SW > $O/prof_dbg1.log 2>&1
LGCN_B200_LIB=$P $SW --k 1 > $O/prof_k1.log 2>&1
LGCN_B200_LIB=$P $SW --d 128 > $O/prof_d128.log 2>&1
tail -n 3 $O/test_tc.log; for f in $O/sweep_*.log $O/prof_*.log; do echo "== $f"; tail -n 3 $f; done
