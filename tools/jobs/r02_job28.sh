#!/bin/bash
# round-2 job 28 (1 GPU): final single-GPU validation — full GPU suite (reference staged), smoke, default bench line, reference arm,
# ncu launch list of the cfg-2 step, ncu --set full of the final top-k kernel
O=gpurun_out/r02ab; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/gputest.log 2>&1; echo "rc=$?" >> $O/gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
timeout 1800 python bench.py > $O/bench_n1.log 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
timeout 900 python bench.py --impl reference > $O/bench_ref.log 2> $O/bench_ref.err; echo "rc=$?" >> $O/bench_ref.err
LEAN="--no-cfg3 --no-cpu-baseline --no-library-bar --no-bf16-block --no-eval --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg2.csv \
  python bench.py --steps 2 --warmup 1 $LEAN > $O/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_topk_tc_kernel -s 1 -c 1 -o $O/tc_topk_full \
  python tools/topk_sweep.py --users 37888 --items 2000000 --reps 1 > $O/ncu_tc.log 2>&1
ls -la $O
tail -n 4 $O/gputest.log $O/smoke.log $O/bench_n1.err $O/bench_ref.err
