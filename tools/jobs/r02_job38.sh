#!/bin/bash
# round-2 job 38 (4 GPUs): cfg-3 block with the graph broadcast from rank 0, auto partition (reduce) — the identity must hold now
O=gpurun_out/r02ak; mkdir -p $O
LEAN="--no-cpu-baseline --no-library-bar --no-bf16-block --no-eval --no-parity"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29581 \
  bench.py --gpus 4 --steps 20 --warmup 5 $LEAN > $O/bench_n4_auto.log 2> $O/bench_n4_auto.err; echo "rc=$?" >> $O/bench_n4_auto.err
tail -n 2 $O/bench_n4_auto.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02ak/bench_n4_auto.log").read().strip().splitlines()[-1])
c = j.get("cfg3", {})
print("auto cfg2x4 ms", j["ms_per_step"], "| cfg3 ms", c.get("ms_per_step"), c.get("parallelism", "")[:50], "parity", c.get("parity"), c.get("error"), "setup_s", c.get("setup_s"))
PY
