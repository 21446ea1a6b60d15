#!/bin/bash
# 1 GPU: L2 locality hints of the column-block backward kernel (EGC_BWD_NEAR_MB = span of target-stream rows around a
# column fetched evict_last, the rest evict_first; 0 = no hints), arxiv- and mag-shaped layers + the uniform graph.
set -x
mkdir -p gpurun_out
T=r02o
for mb in 0 32 64 128; do
  for w in arxiv mag; do
    EGC_BWD_NEAR_MB=$mb timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-extras \
      > gpurun_out/${T}_near${mb}_$w.json 2> gpurun_out/${T}_near${mb}_$w.err
    python - gpurun_out/${T}_near${mb}_$w.json $mb $w <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('near_mb', sys.argv[2], sys.argv[3], 'ms', round(d['ms_per_step'],4), 'scatter', round(d['kernels']['k_scatter_bwd']['ms_per_step'],4))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
  done
done
for mb in 0 64; do
  EGC_BWD_NEAR_MB=$mb timeout 300 python bench.py --locality 0 --steps 20 --warmup 5 --no-cpu-baseline --no-extras \
      > gpurun_out/${T}_near${mb}_uniform.json 2> gpurun_out/${T}_near${mb}_uniform.err
  python - gpurun_out/${T}_near${mb}_uniform.json $mb <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('near_mb', sys.argv[2], 'uniform ms', round(d['ms_per_step'],4), 'scatter', round(d['kernels']['k_scatter_bwd']['ms_per_step'],4))
PY
done
