#!/bin/bash
# round-2 job 35 (4 GPUs): locate the 8e-5 deviation of the A_hat sqrt(deg) identity under the reduce partition at 4 ranks
O=gpurun_out/r02ah; mkdir -p $O
for W in 4 3 2; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29591 \
  tools/debug_reduce_parity.py --shape hbm --partition reduce > $O/hbm_reduce_w$W.log 2>&1
grep "^K=" $O/hbm_reduce_w$W.log | cut -c1-420
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29591 \
  tools/debug_reduce_parity.py --shape cfg3 --partition reduce > $O/cfg3_reduce_w4.log 2>&1
grep "^K=" $O/cfg3_reduce_w4.log | cut -c1-420; tail -n 3 $O/cfg3_reduce_w4.log | cut -c1-300
