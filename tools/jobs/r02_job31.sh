#!/bin/bash
# round-2 job 31 (8 GPUs): cfg-3 strong scaling under the reduce partition vs the default (two-sided) partition, same box;
# the second run is also the N=8 regression of the final build (cfg-2 x 8 headline with its parity object)
O=gpurun_out/r02ae; mkdir -p $O
LEAN="--no-cpu-baseline --no-library-bar --no-bf16-block --no-eval"
for P in reduce auto; do
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 \
  bench.py --gpus 8 --steps 20 --warmup 5 --partition $P $LEAN > $O/bench_n8_$P.log 2> $O/bench_n8_$P.err; echo "rc=$?" >> $O/bench_n8_$P.err
done
tail -n 2 $O/bench_n8_reduce.err $O/bench_n8_auto.err
python - <<'PY'
import json
for P in ("reduce", "auto"):
    try:
        j = json.loads(open(f"gpurun_out/r02ae/bench_n8_{P}.log").read().strip().splitlines()[-1])
        c = j.get("cfg3", {})
        print(P, "cfg2x8 ms", j["ms_per_step"], "parity", j.get("parity", {}).get("ok"), "| cfg3 ms", c.get("ms_per_step"), "parity", c.get("parity"), c.get("error"))
    except Exception as e:
        print(P, "failed", e)
PY
