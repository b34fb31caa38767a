#!/bin/bash
# round-2 job 31 (8 GPUs, budget-safe): cfg-3 strong scaling under the reduce partition (the two-sided number of the
# same code is in profiles/r02_bench_n8_final.log: 66.5 ms/step)
O=gpurun_out/r02ae; mkdir -p $O
LEAN="--no-cpu-baseline --no-library-bar --no-bf16-block --no-eval --no-parity"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 \
  bench.py --gpus 8 --steps 10 --warmup 3 --partition reduce $LEAN > $O/bench_n8_reduce.log 2> $O/bench_n8_reduce.err; echo "rc=$?" >> $O/bench_n8_reduce.err
tail -n 2 $O/bench_n8_reduce.err
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r02ae/bench_n8_reduce.log").read().strip().splitlines()[-1])
    c = j.get("cfg3", {})
    print("reduce cfg2x8 ms", j["ms_per_step"], "| cfg3 ms", c.get("ms_per_step"), "spmm us", c.get("roofline", {}).get("avg_launch_us"), "parity", c.get("parity"), c.get("error"))
except Exception as e:
    print("failed", e)
PY
