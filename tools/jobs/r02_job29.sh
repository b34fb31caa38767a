#!/bin/bash
# round-2 job 29 (1 GPU): early hand-back of the TMEM stage on the last chunk — parity, A-B, phase profile
O=gpurun_out/r02ac; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_cfg2.py tests/test_gpu_parity.py -q -x > $O/test_tc.log 2>&1; echo "rc=$?" >> $O/test_tc.log
for L in g2 m2g4; do LGCN_TC_LAYOUT=$L timeout 600 python -m pytest tests/test_gpu_tc.py -q -x > $O/test_tc_$L.log 2>&1; echo "rc=$?" >> $O/test_tc_$L.log; done
SW="timeout 120 python tools/topk_sweep.py --users 75776 --items 2000000"
P=$PWD/furusato_recommend_b200/liblgcn_b200_tcprof.so
$SW > $O/sweep_default.log 2>&1
$SW --k 1 > $O/sweep_k1.log 2>&1
$SW --k 10 > $O/sweep_k10.log 2>&1
$SW --d 128 > $O/sweep_d128.log 2>&1
$SW --d 128 --k 50 > $O/sweep_d128_k50.log 2>&1
$SW --npos 200 > $O/sweep_npos200.log 2>&1
LGCN_TC_DEBUG=3 $SW > $O/sweep_dbg3.log 2>&1
timeout 300 python tools/topk_sweep.py --users 1000000 --items 2000000 > $O/sweep_1m.log 2>&1
LGCN_B200_LIB=$P $SW > $O/prof_default.log 2>&1
tail -n 3 $O/test_tc*.log; for f in $O/sweep_*.log; do echo "$(basename $f .log): $(tail -n 1 $f)"; done; head -n 14 $O/prof_default.log | cut -c1-300
