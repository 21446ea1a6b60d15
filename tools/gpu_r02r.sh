#!/bin/bash
# 1 GPU: ring variant of the column-block backward kernel (EGC_BWD_RING = ring depth; 0 = register-gather kernel):
# parity tests with the ring on, then the arxiv-shaped bench line (parity_check vs the fp64 oracle included) per depth.
set -x
mkdir -p gpurun_out
T=r02r
EGC_BWD_RING=4 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_epilogue.py -m gpu -q -x > gpurun_out/${T}_pytest_ring4.log 2>&1
echo "pytest ring4 rc=$?"; tail -3 gpurun_out/${T}_pytest_ring4.log
for r in 0 4 7; do
  EGC_BWD_RING=$r timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/${T}_ring${r}.json 2> gpurun_out/${T}_ring${r}.err
  python - gpurun_out/${T}_ring${r}.json $r <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('ring', sys.argv[2], 'ms', round(d['ms_per_step'],4), 'scatter', round(d['kernels']['k_scatter_bwd']['ms_per_step'],4), 'parity', d['parity_check'].get('ok'), d['parity_check'].get('max_rel_err'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
EGC_BWD_RING=4 timeout 200 python bench.py --locality 0 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/${T}_ring4_uniform.json 2> /dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02r_ring4_uniform.json').read().strip().splitlines()[-1])
print('ring 4 uniform ms', round(d['ms_per_step'],4), 'scatter', round(d['kernels']['k_scatter_bwd']['ms_per_step'],4))
PY
