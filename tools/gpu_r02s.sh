#!/bin/bash
# Final single-GPU check of round 2 (r02s), ring kernel on by default: every GPU test file except the 2-rank ones
# (tests/test_gpu_dist.py runs in tools/gpu_r02s_dist.sh on a 2-GPU box) and the default bench line.
set -x
mkdir -p gpurun_out
T=r02s
timeout 600 python -m pytest tests -m gpu -q -rs --ignore tests/test_gpu_dist.py > gpurun_out/${T}_pytest_1gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest_1gpu.log
tail -4 gpurun_out/${T}_pytest_1gpu.log; grep -E "FAILED|ERROR" gpurun_out/${T}_pytest_1gpu.log | head
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02s_bench_1gpu.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['parity_check'].get('ok'), d['roofline']['frac'], d['step_roofline']['frac'])
print({k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
print({k:v['ms_per_step'] for k,v in d['other_configs'].items()})
PY
