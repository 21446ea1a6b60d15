#!/bin/bash
# GPU checkpoint of the whole repo: smoke(), the GPU test suite, the default 1-GPU bench line, the fused-epilogue bench.
# usage: bash tools/gpu_suite.sh <tag> [pytest arguments]      (2-GPU box: the partitioned / data-parallel tests run too;
#        1-GPU box: pass `--ignore tests/test_gpu_dist.py` and run that file alone on a 2-GPU box, as r02s did)
set -x
mkdir -p gpurun_out
T=${1:-suite}
shift
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 900 python -m pytest tests -m gpu -q -rs "$@" > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -8 gpurun_out/${T}_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
head -c 400 gpurun_out/${T}_bench_1gpu.json; echo
timeout 200 python tools/epilogue_bench.py > gpurun_out/${T}_epilogue_bench.json 2> gpurun_out/${T}_epilogue_bench.err
cat gpurun_out/${T}_epilogue_bench.json
