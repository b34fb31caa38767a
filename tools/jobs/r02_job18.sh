#!/bin/bash
# round-2 job 18 (2 GPUs): reduce partition — distributed tests, then bench N=2 with --partition reduce (bf16 block exercises bf16 -> fp32 partials)
O=gpurun_out/r02r; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dist.py -q > $O/test_dist.log 2>&1; echo "rc=$?" >> $O/test_dist.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 \
  bench.py --gpus 2 --steps 20 --warmup 5 --partition reduce --no-eval > $O/bench_n2_reduce.log 2> $O/bench_n2_reduce.err; echo "rc=$?" >> $O/bench_n2_reduce.err
tail -n 5 $O/test_dist.log; tail -n 5 $O/bench_n2_reduce.err; tail -c 800 $O/bench_n2_reduce.log
