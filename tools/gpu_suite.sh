#!/bin/bash
# The whole GPU test suite (2 GPUs: the partitioned / data-parallel tests run too) + the default 1-GPU bench line.
# usage: bash tools/gpu_suite.sh <tag>
set -x
mkdir -p gpurun_out
T=${1:-suite}
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -8 gpurun_out/${T}_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
head -c 400 gpurun_out/${T}_bench_1gpu.json
