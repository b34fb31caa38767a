#!/bin/bash
# round-2 job 23 (1 GPU): where a candidate-path entry spends its ~1600 clocks
O=gpurun_out/r02w; mkdir -p $O
SW="timeout 120 python tools/topk_sweep.py --users 75776 --items 2000000"
export LGCN_B200_LIB=$PWD/furusato_recommend_b200/liblgcn_b200_tcprof.so
$SW > $O/prof_m2g2.log 2>&1
$SW --k 1 > $O/prof_m2g2_k1.log 2>&1
$SW --npos 0 > $O/prof_m2g2_npos0.log 2>&1
for f in $O/prof_*.log; do echo "== $f"; head -n 4 $f | cut -c1-400; done
