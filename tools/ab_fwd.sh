#!/bin/bash
# A/B of the forward aggregation kernels (row-block vs warp-per-row) + parity of the new one.  Run under gpurun.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
summ() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],4), {k: round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
" $1; }
for wl in arxiv mag; do
  python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_${wl}_rows.json 2>gpurun_out/ab_${wl}_rows.err
  summ gpurun_out/ab_${wl}_rows.json
  EGC_FWD_WARP_PER_ROW=1 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_${wl}_wpr.json 2>gpurun_out/ab_${wl}_wpr.err
  summ gpurun_out/ab_${wl}_wpr.json
done
