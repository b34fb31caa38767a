#!/bin/bash
# round-2 job 42 (1 GPU): HEAD's bench line with the cfg-3 block (host-built popularity CDF), lean
O=gpurun_out/r02ao; mkdir -p $O
timeout 150 python bench.py --no-cpu-baseline --no-library-bar --no-bf16-block > $O/bench_n1_lean.log 2> $O/bench_n1_lean.err; echo "rc=$?" >> $O/bench_n1_lean.err
tail -n 1 $O/bench_n1_lean.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02ao/bench_n1_lean.log").read().strip().splitlines()[-1])
c = j["cfg3"]
print("ms", j["ms_per_step"], "parity", j["parity"]["ok"], "sweep TF", j["eval"]["sweep_cfg5"]["tflops"], "| cfg3 ms", c["ms_per_step"], "spmm us", c["roofline"]["avg_launch_us"], "frac", c["roofline"]["frac"], "parity", c["parity"]["max_rel_err"], c["workload"][-30:])
PY
