#!/bin/bash
# round-2 job 5 (2 GPUs): distributed tests + the default bench line at N=2 (cfg-2 x2 weak, bf16 block, cfg-3 strong)
O=gpurun_out/r02e; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dist.py -q > $O/test_dist.log 2>&1; echo "rc=$?" >> $O/test_dist.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.log 2> $O/bench_n2.err; echo "rc=$?" >> $O/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --workload cfg5 --eval-users 2000000 > $O/bench_cfg5_n2.log 2> $O/bench_cfg5_n2.err
tail -n 3 $O/test_dist.log $O/bench_n2.err; tail -c 1500 $O/bench_n2.log
