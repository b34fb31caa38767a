#!/bin/bash
# round-2 job 40 (1 GPU): last sanity of HEAD — full GPU suite, smoke, one sweep
O=gpurun_out/r02am; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q > $O/gputest.log 2>&1; echo "rc=$?" >> $O/gputest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
timeout 100 python tools/topk_sweep.py --users 75776 --items 2000000 > $O/sweep.log 2>&1
tail -n 3 $O/gputest.log; tail -n 2 $O/smoke.log | cut -c1-200; tail -n 1 $O/sweep.log
