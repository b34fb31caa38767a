#!/bin/bash
# round-2 job 7 (8 GPUs): the default bench line at N=8 (cfg-2 x8 weak + bf16 block + cfg-3 strong + parity) and cfg-5 at 10 M users
O=gpurun_out/r02m; mkdir -p $O
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 \
  bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.log 2> $O/bench_n8.err; echo "rc=$?" >> $O/bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 \
  bench.py --gpus 8 --workload cfg5 --eval-users 10000000 > $O/bench_cfg5_n8.log 2> $O/bench_cfg5_n8.err; echo "rc=$?" >> $O/bench_cfg5_n8.err
tail -n 3 $O/bench_n8.err $O/bench_cfg5_n8.err; tail -c 1200 $O/bench_n8.log
