#!/bin/bash
# 2-GPU call: whole GPU suite (distributed parity included) + partitioned bench on both workloads
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/c3_pytest.log
tail -3 gpurun_out/c3_pytest.log
for wl in arxiv mag; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --workload $wl > gpurun_out/c3_bench_2gpu_$wl.json 2> gpurun_out/c3_bench_2gpu_$wl.err
  echo "bench $wl rc=$?"; tail -c 1200 gpurun_out/c3_bench_2gpu_$wl.json | head -c 300; echo
done
