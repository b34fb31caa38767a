#!/bin/bash
# round-2 job 16 (1 GPU): the driver's GPU test command on the final tree + smoke
O=gpurun_out/r02p; mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu > $O/gputest.log 2>&1; echo "rc=$?" >> $O/gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
tail -n 5 $O/gputest.log $O/smoke.log
