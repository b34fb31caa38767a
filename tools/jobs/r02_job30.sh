#!/bin/bash
# round-2 job 30 (1 GPU): FINAL single-GPU validation of the shipped build — full GPU suite (reference staged), smoke, sweeps,
# default bench line, reference arm, ncu launch list of the cfg-2 step, ncu --set full of the top-k kernel
O=gpurun_out/r02ad; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/gputest.log 2>&1; echo "rc=$?" >> $O/gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
SW="timeout 120 python tools/topk_sweep.py --users 75776 --items 2000000"
$SW > $O/sweep_default.log 2>&1
$SW --d 128 > $O/sweep_d128.log 2>&1
LGCN_TC_DEBUG=1 $SW > $O/sweep_dbg1.log 2>&1
LGCN_TC_DEBUG=3 $SW > $O/sweep_dbg3.log 2>&1
timeout 300 python tools/topk_sweep.py --users 1000000 --items 2000000 > $O/sweep_1m.log 2>&1
timeout 1800 python bench.py > $O/bench_n1.log 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
timeout 900 python bench.py --impl reference > $O/bench_ref.log 2> $O/bench_ref.err; echo "rc=$?" >> $O/bench_ref.err
LEAN="--no-cfg3 --no-cpu-baseline --no-library-bar --no-bf16-block --no-eval --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg2.csv \
  python bench.py --steps 2 --warmup 1 $LEAN > $O/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_topk_tc_kernel -s 1 -c 1 -o $O/tc_topk_full \
  python tools/topk_sweep.py --users 37888 --items 2000000 --reps 1 > $O/ncu_tc.log 2>&1
for f in $O/sweep_*.log; do echo "$(basename $f .log): $(tail -n 1 $f)"; done
tail -n 4 $O/gputest.log $O/smoke.log $O/bench_n1.err $O/bench_ref.err
