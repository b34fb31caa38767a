#!/bin/bash
# round-2 job 11 (1 GPU): full suite on the final tree, smoke, default bench line, cfg-4 (reference arithmetic and true softmax)
O=gpurun_out/r02k; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/gputest.log 2>&1; echo "rc=$?" >> $O/gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
timeout 1500 python bench.py > $O/bench_n1.log 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
timeout 600 python bench.py --impl reference --steps 30 --warmup 5 > $O/bench_ref.log 2> $O/bench_ref.err
timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 3 > $O/bench_cfg4.log 2> $O/bench_cfg4.err; echo "rc=$?" >> $O/bench_cfg4.err
timeout 600 python bench.py --workload cfg4 --ssm-softmax --steps 3 --warmup 3 > $O/bench_cfg4_ssm.log 2> $O/bench_cfg4_ssm.err; echo "rc=$?" >> $O/bench_cfg4_ssm.err
tail -n 3 $O/gputest.log $O/smoke.log $O/bench_n1.err $O/bench_cfg4.err $O/bench_cfg4_ssm.err
