#!/bin/bash
# A/B of the backward CSC pass (column-block vs warp-per-column) + parity.  Run under gpurun.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
summ() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],4), {k: round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
" $1; }
for wl in arxiv mag; do
  python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_${wl}_cols.json 2>gpurun_out/ab_${wl}_cols.err
  summ gpurun_out/ab_${wl}_cols.json
  EGC_BWD_WARP_PER_COLUMN=1 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_${wl}_wpc.json 2>gpurun_out/ab_${wl}_wpc.err
  summ gpurun_out/ab_${wl}_wpc.json
done
python bench.py --workload mag --steps 20 --warmup 5 --no-cpu-baseline --bwd-flags 64 > gpurun_out/ab_mag_cols_noslab.json 2>gpurun_out/ab_mag_cols_noslab.err
summ gpurun_out/ab_mag_cols_noslab.json
