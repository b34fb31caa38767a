#!/bin/bash
# round-2 job 6 (1 GPU): top-k epilogue experiments after the MMA issue fix
O=gpurun_out/r02f; mkdir -p $O
RS=$PWD/furusato_recommend_b200/liblgcn_b200_rs.so
timeout 300 python -m pytest tests/test_gpu_tc.py -q > $O/tc_test.log 2>&1; echo "rc=$?" >> $O/tc_test.log
LGCN_B200_LIB=$RS LGCN_TC_LAYOUT=m2rs timeout 300 python -m pytest tests/test_gpu_tc.py -q > $O/tc_test_rs.log 2>&1; echo "rc=$?" >> $O/tc_test_rs.log
SW="timeout 120 python tools/topk_sweep.py --users 75776 --items 2000000"
$SW > $O/sweep_default.log 2>&1
LGCN_TC_DEBUG=3 $SW > $O/sweep_default_dbg3.log 2>&1
LGCN_TC_DEBUG=2 $SW > $O/sweep_default_dbg2.log 2>&1
LGCN_B200_LIB=$RS LGCN_TC_LAYOUT=m2rs $SW > $O/sweep_rs.log 2>&1
LGCN_B200_LIB=$RS LGCN_TC_LAYOUT=m2rs LGCN_TC_DEBUG=3 $SW > $O/sweep_rs_dbg3.log 2>&1
for T in 22 24 30 36; do LGCN_TC_TRIG=$T $SW > $O/sweep_trig$T.log 2>&1; done
LGCN_B200_LIB=$RS LGCN_TC_LAYOUT=m2rs LGCN_TC_TRIG=30 $SW > $O/sweep_rs_trig30.log 2>&1
tail -n 2 $O/*.log
