#!/bin/bash
# round-2 job 17 (1 GPU): what does the positives walk cost in the top-k slow path?
O=gpurun_out/r02q; mkdir -p $O
SW="timeout 120 python tools/topk_sweep.py --users 75776 --items 2000000"
$SW > $O/sweep_npos50.log 2>&1
$SW --npos 0 > $O/sweep_npos0.log 2>&1
$SW --npos 5 > $O/sweep_npos5.log 2>&1
$SW --npos 200 > $O/sweep_npos200.log 2>&1
$SW --k 10 > $O/sweep_k10.log 2>&1
$SW --k 5 > $O/sweep_k5.log 2>&1
$SW --k 1 > $O/sweep_k1.log 2>&1
tail -n 1 $O/*.log
