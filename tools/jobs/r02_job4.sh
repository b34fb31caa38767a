#!/bin/bash
# round-2 job 4 (1 GPU): top-k with warp-uniform two-issuer MMA loop, SSM kernel tests
O=gpurun_out/r02d; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ssm.py tests/test_gpu_parity.py -q -k "tc or ssm or topk or eval" > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
LGCN_TC_LAYOUT=m2c2 timeout 300 python -m pytest tests/test_gpu_tc.py -q > $O/tc_test_m2c2.log 2>&1; echo "rc=$?" >> $O/tc_test_m2c2.log
SW="timeout 120 python tools/topk_sweep.py --users 75776 --items 2000000"
$SW > $O/sweep_default.log 2>&1
LGCN_TC_DEBUG=1 $SW > $O/sweep_default_dbg1.log 2>&1
LGCN_TC_DEBUG=3 $SW > $O/sweep_default_dbg3.log 2>&1
LGCN_TC_LAYOUT=m2c2 $SW > $O/sweep_m2c2.log 2>&1
LGCN_TC_LAYOUT=m2c2 LGCN_TC_DEBUG=1 $SW > $O/sweep_m2c2_dbg1.log 2>&1
LGCN_TC_LAYOUT=g2 $SW > $O/sweep_g2.log 2>&1
LGCN_TC_LAYOUT=m2g4 $SW > $O/sweep_m2g4.log 2>&1
$SW --d 128 > $O/sweep_d128.log 2>&1
$SW --d 128 --k 50 > $O/sweep_d128_k50.log 2>&1
timeout 300 python tools/topk_sweep.py --users 1000000 --items 2000000 > $O/sweep_1m.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_topk_tc_kernel -s 1 -c 1 -o $O/tc_topk_full \
  python tools/topk_sweep.py --users 37888 --items 2000000 --reps 1 > $O/ncu_tc.log 2>&1
tail -n 2 $O/*.log
