#!/bin/bash
# round-2 job 2 (1 GPU): full GPU suite, default bench line, ncu launch list + full captures (SpMM on cfg-3, top-k)
O=gpurun_out/r02b; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/gputest.log 2>&1; echo "rc=$?" >> $O/gputest.log
timeout 1500 python bench.py > $O/bench_n1.log 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_ref.log 2> $O/bench_ref.err
LEAN="--no-cfg3 --no-cpu-baseline --no-library-bar --no-bf16-block --no-eval --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg2.csv \
  python bench.py --steps 2 --warmup 1 $LEAN > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:spmm_layer_kernel -s 6 -c 6 -o $O/spmm_cfg3_full \
  python bench.py --workload cfg3 --steps 1 --warmup 3 > $O/ncu_spmm_cfg3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_topk_tc_kernel -s 1 -c 1 -o $O/tc_topk_full \
  python tools/topk_sweep.py --users 37888 --items 2000000 --reps 1 > $O/ncu_tc.log 2>&1
ls -la $O
tail -3 $O/gputest.log $O/bench_n1.err
