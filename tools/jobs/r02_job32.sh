#!/bin/bash
# round-2 job 32 (4 GPUs): cfg-3 strong scaling under the reduce partition vs the default (two-sided) partition, same box;
# the second run is also the N=4 regression of the final build (cfg-2 x 4 headline with its parity object)
O=gpurun_out/r02af; mkdir -p $O
LEAN="--no-cpu-baseline --no-library-bar --no-bf16-block --no-eval"
for P in reduce auto; do
timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29581 \
  bench.py --gpus 4 --steps 20 --warmup 5 --partition $P $LEAN > $O/bench_n4_$P.log 2> $O/bench_n4_$P.err; echo "rc=$?" >> $O/bench_n4_$P.err
done
tail -n 2 $O/bench_n4_reduce.err $O/bench_n4_auto.err
python - <<'PY'
import json
for P in ("reduce", "auto"):
    try:
        j = json.loads(open(f"gpurun_out/r02af/bench_n4_{P}.log").read().strip().splitlines()[-1])
        c = j.get("cfg3", {})
        print(P, "cfg2x4 ms", j["ms_per_step"], "parity", j.get("parity", {}).get("ok"), "| cfg3 ms", c.get("ms_per_step"), "spmm us", c.get("roofline", {}).get("avg_launch_us"), "parity", c.get("parity"), c.get("error"))
    except Exception as e:
        print(P, "failed", e)
PY
