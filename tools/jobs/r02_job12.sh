#!/bin/bash
# round-2 job 9 (2 GPUs): regression of the distributed tests + bench N=2 with interleaved row order and the cfg-3 bf16 sub-block
O=gpurun_out/r02l; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dist.py -q > $O/test_dist.log 2>&1; echo "rc=$?" >> $O/test_dist.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
  bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.log 2> $O/bench_n2.err; echo "rc=$?" >> $O/bench_n2.err
tail -n 3 $O/test_dist.log $O/bench_n2.err; tail -c 600 $O/bench_n2.log
