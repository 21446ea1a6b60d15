#!/bin/bash
set -x
mkdir -p gpurun_out
N=${1:-2}
for c in 1 2 4 8 16; do
  for w in 128 64; do
    EGC_PEER_PUSH_CTAS_PER_SM=$c timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 \
      tools/peer_bw.py $w 64 2>&1 | grep -E "GB/s|Error|error" >> gpurun_out/peer_bw_${N}gpu.txt
  done
done
cat gpurun_out/peer_bw_${N}gpu.txt
