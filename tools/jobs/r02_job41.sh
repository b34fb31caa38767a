#!/bin/bash
# round-2 job 41 (1 GPU): the opt-in experiment layouts of the top-k kernel still pass the bit-exact suite with the final candidate path
O=gpurun_out/r02an; mkdir -p $O
for L in m2s4 m2c2 m2c4 m2rl; do LGCN_TC_LAYOUT=$L timeout 200 python -m pytest tests/test_gpu_tc.py -q -x > $O/test_tc_$L.log 2>&1; echo "$L rc=$? $(tail -n 1 $O/test_tc_$L.log)"; done
