#!/bin/bash
# Final state of round 2 on a 2-GPU box (r02q): the whole GPU suite (partitioned / data-parallel tests included) and the
# default single-GPU bench line.
set -x
mkdir -p gpurun_out
T=r02q
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -6 gpurun_out/${T}_pytest.log
grep -E "FAILED|ERROR" gpurun_out/${T}_pytest.log | head
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02q_bench_1gpu.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['parity_check'].get('ok'), d['roofline']['frac'], d['step_roofline']['frac'])
print({k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
print({k:v['ms_per_step'] for k,v in d['other_configs'].items()})
PY
