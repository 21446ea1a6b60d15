#!/bin/bash
# Final 2-GPU check of round 2 (r02s): the partitioned / data-parallel tests with the ring kernel on by default.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -rs > gpurun_out/r02s_pytest_dist_2gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02s_pytest_dist_2gpu.log
tail -4 gpurun_out/r02s_pytest_dist_2gpu.log; grep -E "FAILED|ERROR" gpurun_out/r02s_pytest_dist_2gpu.log | head
