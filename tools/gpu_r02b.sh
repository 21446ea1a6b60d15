#!/bin/bash
# Round 2, GPU call B (2 GPUs): test suite after the backward column phases / fused all-reduce / deterministic routing,
# 1-GPU and 2-GPU bench lines.
set -x
mkdir -p gpurun_out
T=${1:-r02b}
timeout 900 python -m pytest tests -m gpu -q -rs --durations=5 -x > gpurun_out/${T}_pytest_x.log 2>&1
echo "pytest -x rc=$?" >> gpurun_out/${T}_pytest_x.log
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_stack.py tests/test_gpu_dist.py -m gpu -q -rs > gpurun_out/${T}_pytest_new.log 2>&1
echo "pytest new rc=$?" >> gpurun_out/${T}_pytest_new.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err
echo "bench2 rc=$?"
tail -5 gpurun_out/${T}_pytest_x.log gpurun_out/${T}_pytest_new.log
tail -c 400 gpurun_out/${T}_bench_2gpu.err
head -c 400 gpurun_out/${T}_bench_2gpu.json
