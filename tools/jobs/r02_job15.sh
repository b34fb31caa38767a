#!/bin/bash
# round-2 job 15 (1 GPU): f16 scorer with the new issue loop; ncu --set full of the cfg-2 SpMM launches of the final build
O=gpurun_out/r02o; mkdir -p $O
SW="timeout 120 python tools/topk_sweep.py --users 75776 --items 2000000"
$SW --precision f16 > $O/sweep_f16.log 2>&1
LGCN_TC_DEBUG=3 $SW --precision f16 > $O/sweep_f16_dbg3.log 2>&1
LGCN_TC_DEBUG=1 $SW --precision f16 > $O/sweep_f16_dbg1.log 2>&1
$SW > $O/sweep_bf16.log 2>&1
LEAN="--no-cfg3 --no-cpu-baseline --no-library-bar --no-bf16-block --no-eval --no-parity"
timeout 600 ncu --set full --clock-control none -k regex:spmm_layer_kernel -s 12 -c 6 -o $O/spmm_cfg2_full \
  python bench.py --steps 2 --warmup 3 $LEAN > $O/ncu_spmm_cfg2.log 2>&1
tail -n 2 $O/sweep_*.log
