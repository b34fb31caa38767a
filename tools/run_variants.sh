#!/bin/bash
# Tuning aid: time every liblgcn_b200_*.so variant at cfg-2 (run on the GPU box through gpurun).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/variants.log
: > $out
for lib in furusato_recommend_b200/liblgcn_b200.so furusato_recommend_b200/liblgcn_b200_*.so; do
  for st in ${STORAGES:-fp32 bf16}; do
    LGCN_B200_LIB=$PWD/$lib STORAGE=$st D=${D:-64} timeout 120 python tools/spmm_bench.py 2>&1 | grep -v Warning | tail -1 >> $out
  done
done
cat $out
