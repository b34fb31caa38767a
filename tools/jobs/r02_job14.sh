#!/bin/bash
# round-2 job 14 (1 GPU): robustness tests + full suite with -x (the driver's command) + quick bench for the e2e figure
O=gpurun_out/r02n; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_robustness.py -q > $O/robust.log 2>&1; echo "rc=$?" >> $O/robust.log
timeout 1200 python -m pytest tests -x -q -m gpu > $O/gputest.log 2>&1; echo "rc=$?" >> $O/gputest.log
timeout 600 python bench.py --no-cfg3 --no-cpu-baseline --no-library-bar --no-bf16-block > $O/bench_quick.log 2> $O/bench_quick.err; echo "rc=$?" >> $O/bench_quick.err
tail -n 4 $O/robust.log $O/gputest.log $O/bench_quick.err
