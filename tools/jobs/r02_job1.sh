#!/bin/bash
# round-2 job 1 (1 GPU): sanity, top-k A-B (3-input max tree; register-staged epilogue), sanitizer, set-up timing
O=gpurun_out/r02a; mkdir -p $O
RS=$PWD/furusato_recommend_b200/liblgcn_b200_rs.so
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputest.log 2>&1; echo "rc=$?" >> $O/gputest.log
SW="python tools/topk_sweep.py --users 75776 --items 2000000"
timeout 120 $SW > $O/sweep_default.log 2>&1
for dbg in 1 2 3; do LGCN_TC_DEBUG=$dbg timeout 120 $SW > $O/sweep_debug$dbg.log 2>&1; done
LGCN_B200_LIB=$RS LGCN_TC_LAYOUT=m2rs timeout 300 python -m pytest tests/test_gpu_tc.py -x -q > $O/rs_test.log 2>&1; echo "rc=$?" >> $O/rs_test.log
LGCN_B200_LIB=$RS LGCN_TC_LAYOUT=m2rs timeout 120 $SW > $O/sweep_rs.log 2>&1
timeout 120 $SW --d 128 > $O/sweep_d128.log 2>&1
timeout 600 python tools/build_timing.py > $O/build_timing_cfg3.log 2>&1
timeout 200 python tools/build_timing.py 2400000 600000 75000000 > $O/build_timing_hbm.log 2>&1
timeout 500 compute-sanitizer --tool memcheck --log-file $O/memcheck_tc.txt python -m pytest tests/test_gpu_tc.py -x -q > $O/memcheck_tc.log 2>&1; echo "rc=$?" >> $O/memcheck_tc.log
timeout 500 compute-sanitizer --tool memcheck --log-file $O/memcheck_hub.txt python -m pytest tests/test_gpu_parity.py -x -q -k "hub_rows or stage_one or sampler_bit" > $O/memcheck_hub.log 2>&1; echo "rc=$?" >> $O/memcheck_hub.log
timeout 500 compute-sanitizer --tool racecheck --log-file $O/racecheck_hub.txt python -m pytest tests/test_gpu_parity.py -x -q -k "hub_rows and 64" > $O/racecheck_hub.log 2>&1; echo "rc=$?" >> $O/racecheck_hub.log
timeout 500 compute-sanitizer --tool racecheck --log-file $O/racecheck_tc.txt python -m pytest tests/test_gpu_tc.py -x -q -k "many_tiles" > $O/racecheck_tc.log 2>&1; echo "rc=$?" >> $O/racecheck_tc.log
tail -3 $O/*.log
