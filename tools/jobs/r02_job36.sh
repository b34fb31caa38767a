#!/bin/bash
# round-2 job 36 (4 GPUs): the identity under the reduce partition, fresh model vs after train steps (hbm shape, then cfg-3)
O=gpurun_out/r02ai; mkdir -p $O
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29591 \
  tools/debug_reduce_parity.py --shape hbm --partition reduce > $O/hbm_reduce_w4.log 2>&1
grep "^K=3" $O/hbm_reduce_w4.log | cut -c1-330; tail -n 2 $O/hbm_reduce_w4.log | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29591 \
  tools/debug_reduce_parity.py --shape cfg3 --partition reduce > $O/cfg3_reduce_w4.log 2>&1
grep "^K=3" $O/cfg3_reduce_w4.log | cut -c1-330; tail -n 2 $O/cfg3_reduce_w4.log | cut -c1-200
