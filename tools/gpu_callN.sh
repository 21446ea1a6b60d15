#!/bin/bash
# N-GPU call: partitioned bench on both workloads (N = $1)
N=${1:-4}
mkdir -p gpurun_out
for wl in arxiv mag; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --workload $wl > gpurun_out/cN_bench_${N}gpu_$wl.json 2> gpurun_out/cN_bench_${N}gpu_$wl.err
  echo "bench $wl rc=$?"; tail -3 gpurun_out/cN_bench_${N}gpu_$wl.err | cut -c1-300
  python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d['n_gpus'], 'ms', round(d['ms_per_step'],4), 'G edges/s', round(d['value']/1e9,3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'halo rows', d['config']['halo_rows_total'])
print({k: round(v['ms_per_step'],4) for k,v in d['kernels_rank0'].items()})
" gpurun_out/cN_bench_${N}gpu_$wl.json
done
