#!/bin/bash
# round-2 job 3 (1 GPU): full suite incl. boundary + ingest, top-k tile-supply experiments, quick bench
O=gpurun_out/r02c; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/gputest.log 2>&1; echo "rc=$?" >> $O/gputest.log
for L in m2c2 m2c4; do
  LGCN_TC_LAYOUT=$L timeout 300 python -m pytest tests/test_gpu_tc.py -q > $O/tc_test_$L.log 2>&1; echo "rc=$?" >> $O/tc_test_$L.log
done
SW="timeout 120 python tools/topk_sweep.py --users 75776 --items 2000000"
$SW > $O/sweep_default.log 2>&1
LGCN_TC_DEBUG=1 $SW > $O/sweep_default_dbg1.log 2>&1
LGCN_TC_SPLIT=4 $SW > $O/sweep_split4.log 2>&1
LGCN_TC_SPLIT=4 LGCN_TC_DEBUG=1 $SW > $O/sweep_split4_dbg1.log 2>&1
LGCN_TC_SPLIT=8 LGCN_TC_DEBUG=1 $SW > $O/sweep_split8_dbg1.log 2>&1
LGCN_TC_STAGES=3 LGCN_TC_DEBUG=1 $SW > $O/sweep_stages3_dbg1.log 2>&1
LGCN_TC_STAGES=3 $SW > $O/sweep_stages3.log 2>&1
for L in m2c2 m2c4; do
  LGCN_TC_LAYOUT=$L $SW > $O/sweep_$L.log 2>&1
  LGCN_TC_LAYOUT=$L LGCN_TC_DEBUG=1 $SW > $O/sweep_${L}_dbg1.log 2>&1
  LGCN_TC_LAYOUT=$L LGCN_TC_DEBUG=3 $SW > $O/sweep_${L}_dbg3.log 2>&1
done
LGCN_TC_DEBUG=3 $SW > $O/sweep_default_dbg3.log 2>&1
$SW --d 128 > $O/sweep_d128.log 2>&1
timeout 900 python bench.py --no-cfg3 --steps 10 --no-bf16-block > $O/bench_quick.log 2> $O/bench_quick.err; echo "rc=$?" >> $O/bench_quick.err
tail -n 2 $O/*.log
