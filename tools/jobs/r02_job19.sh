#!/bin/bash
# round-2 job 19 (2 GPUs): reduce partition — distributed tests, then bench N=2 A-B (--partition reduce vs auto), no cfg-3 block
O=gpurun_out/r02s; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dist.py -q > $O/test_dist.log 2>&1; echo "rc=$?" >> $O/test_dist.log
for P in reduce auto; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 \
  bench.py --gpus 2 --steps 30 --warmup 5 --partition $P --no-eval --no-cfg3 > $O/bench_n2_$P.log 2> $O/bench_n2_$P.err; echo "rc=$?" >> $O/bench_n2_$P.err
done
tail -n 5 $O/test_dist.log; tail -n 3 $O/bench_n2_*.err; tail -c 600 $O/bench_n2_reduce.log; tail -c 600 $O/bench_n2_auto.log
