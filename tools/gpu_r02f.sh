#!/bin/bash
# Round 2, GPU call F (1 GPU): full GPU suite after the MN-major wgrad fix, 1-GPU bench, then the ncu evidence pass.
set -x
mkdir -p gpurun_out
T=${1:-r02f}
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -8 gpurun_out/${T}_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
bash tools/gpu_r02_ncu.sh $T
