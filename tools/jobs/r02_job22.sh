#!/bin/bash
# round-2 job 22 (1 GPU): pipeline timeline (clock64 stamps) of the top-k kernel, m2g2 and m2rl, full / dbg3 / dbg1
O=gpurun_out/r02v; mkdir -p $O
SW="timeout 120 python tools/topk_sweep.py --users 75776 --items 2000000"
export LGCN_B200_LIB=$PWD/furusato_recommend_b200/liblgcn_b200_tcprof.so
$SW > $O/prof_m2g2.log 2>&1
LGCN_TC_DEBUG=3 $SW > $O/prof_m2g2_dbg3.log 2>&1
LGCN_TC_DEBUG=1 $SW > $O/prof_m2g2_dbg1.log 2>&1
LGCN_TC_LAYOUT=m2rl $SW > $O/prof_m2rl.log 2>&1
LGCN_TC_LAYOUT=m2rl LGCN_TC_DEBUG=3 $SW > $O/prof_m2rl_dbg3.log 2>&1
LGCN_TC_LAYOUT=m2rl LGCN_TC_DEBUG=2 $SW > $O/prof_m2rl_dbg2.log 2>&1
LGCN_TC_LAYOUT=m2rl LGCN_TC_DEBUG=1 $SW > $O/prof_m2rl_dbg1.log 2>&1
for f in $O/prof_*.log; do echo "== $f"; tail -n 12 $f; done
