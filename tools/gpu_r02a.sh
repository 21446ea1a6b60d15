#!/bin/bash
# Round 2, GPU call A (2 GPUs): whole GPU test suite incl. the 2-rank partitioned / data-parallel tests, 1-GPU bench
# (default line with other_configs), persisting-L2 A/B runs, uniform-graph run, 2-GPU bench with parity_check.
set -x
mkdir -p gpurun_out
T=r02a
nvidia-smi -L > gpurun_out/${T}_gpus.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -rs --durations=8 > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
echo "bench1 rc=$?"
for m in 1 2 3; do
  EGC_L2_PERSIST=$m timeout 200 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/${T}_persist$m.json 2> gpurun_out/${T}_persist$m.err
done
timeout 200 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --locality 0 > gpurun_out/${T}_uniform.json 2> gpurun_out/${T}_uniform.err
EGC_L2_PERSIST=3 timeout 200 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --locality 0 > gpurun_out/${T}_uniform_persist3.json 2> gpurun_out/${T}_uniform_persist3.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err
echo "bench2 rc=$?"
tail -5 gpurun_out/${T}_pytest.log
for f in gpurun_out/${T}_*.json; do echo "== $f"; head -c 600 $f; echo; done
tail -3 gpurun_out/${T}_bench_2gpu.err gpurun_out/${T}_bench_1gpu.err
