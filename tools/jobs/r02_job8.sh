#!/bin/bash
# round-2 job 8 (1 GPU): accumulator hand-off experiments (wait hints, 4 TMEM stages)
O=gpurun_out/r02h; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_tc.py -q > $O/tc_test.log 2>&1; echo "rc=$?" >> $O/tc_test.log
LGCN_TC_LAYOUT=m2s4 timeout 300 python -m pytest tests/test_gpu_tc.py -q > $O/tc_test_m2s4.log 2>&1; echo "rc=$?" >> $O/tc_test_m2s4.log
SW="timeout 120 python tools/topk_sweep.py --users 75776 --items 2000000"
$SW > $O/sweep_default.log 2>&1
LGCN_TC_DEBUG=3 $SW > $O/sweep_default_dbg3.log 2>&1
for H in 0 200 2000 20000; do
  LGCN_TC_WAIT_NS=$H $SW > $O/sweep_wait$H.log 2>&1
  LGCN_TC_WAIT_NS=$H LGCN_TC_DEBUG=3 $SW > $O/sweep_wait${H}_dbg3.log 2>&1
done
LGCN_TC_LAYOUT=m2s4 $SW > $O/sweep_m2s4.log 2>&1
LGCN_TC_LAYOUT=m2s4 LGCN_TC_DEBUG=3 $SW > $O/sweep_m2s4_dbg3.log 2>&1
LGCN_TC_LAYOUT=m2s4 LGCN_TC_DEBUG=1 $SW > $O/sweep_m2s4_dbg1.log 2>&1
tail -n 2 $O/*.log
