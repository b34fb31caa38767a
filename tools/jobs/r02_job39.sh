#!/bin/bash
# round-2 job 39 (N GPUs, N = $1): the default bench line of the final build (what the driver's scaling run launches) + dist tests at N = 2
N=$1
O=gpurun_out/r02al_n$N; mkdir -p $O
if [ "$N" = "2" ]; then timeout 400 python -m pytest tests/test_gpu_dist.py -q > $O/test_dist.log 2>&1; echo "rc=$?" >> $O/test_dist.log; tail -n 3 $O/test_dist.log; fi
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 \
  bench.py --gpus $N --steps 30 --warmup 5 > $O/bench.log 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
tail -n 2 $O/bench.err
python - <<PY
import json
j = json.loads(open("$O/bench.log").read().strip().splitlines()[-1])
c = j.get("cfg3", {})
print("N=$N value", j["value"], "ms", j["ms_per_step"], "parity", j.get("parity", {}).get("ok"), "partition", j["config"].get("parallelism", "")[:40])
print("  bf16_storage", {k: j.get("bf16_storage", {}).get(k) for k in ("ms_per_step",)}, (j.get("bf16_storage", {}).get("parity") or {}).get("ok"))
print("  cfg3 ms", c.get("ms_per_step"), c.get("parallelism", "")[:50], "parity", (c.get("parity") or {}).get("ok"), c.get("error"))
print("  cfg3 bf16", (c.get("bf16_storage") or {}).get("ms_per_step"), ((c.get("bf16_storage") or {}).get("parity") or {}))
e = j.get("eval") or {}
print("  eval sweep", (e.get("sweep_cfg5") or {}).get("value"), (e.get("sweep_cfg5") or {}).get("tflops_per_gpu"))
PY
