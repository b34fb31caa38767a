#!/bin/bash
# round-2 job 34 (N GPUs, N = $1): cfg-3 strong scaling, auto partition (reduce for the user-heavy graph) vs two_sided, same box
N=$1
O=gpurun_out/r02ag_n$N; mkdir -p $O
LEAN="--no-cpu-baseline --no-library-bar --no-bf16-block --no-eval"
for P in auto two_sided; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 \
  bench.py --gpus $N --steps 20 --warmup 5 --partition $P $LEAN > $O/bench_$P.log 2> $O/bench_$P.err; echo "rc=$?" >> $O/bench_$P.err
done
tail -n 2 $O/bench_auto.err $O/bench_two_sided.err
python - <<PY
import json
for P in ("auto", "two_sided"):
    try:
        j = json.loads(open("$O/bench_%s.log" % P).read().strip().splitlines()[-1])
        c = j.get("cfg3", {})
        print(P, "cfg2xN ms", j["ms_per_step"], "parity", j.get("parity", {}).get("ok"), "| cfg3 ms", c.get("ms_per_step"), c.get("parallelism", "")[:60], "parity", c.get("parity", {}).get("max_rel_err"), c.get("error"))
    except Exception as e:
        print(P, "failed", e)
PY
