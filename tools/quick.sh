#!/bin/bash
# parity suite + both 1-GPU bench workloads, per-kernel summary
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for wl in arxiv mag; do
  python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/q_${wl}.json 2>gpurun_out/q_${wl}.err
  python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],4), {k: round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
" gpurun_out/q_${wl}.json
done
