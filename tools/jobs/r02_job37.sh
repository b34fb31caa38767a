#!/bin/bash
# round-2 job 37 (2 GPUs): is the synthetic cfg-3 graph bit-identical on every rank / every call?
O=gpurun_out/r02aj; mkdir -p $O
timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 \
  tools/debug_graph_determinism.py cfg3 > $O/det_cfg3.log 2>&1
grep -v Warning $O/det_cfg3.log | tail -n 8
