#!/bin/bash
# 8 GPUs, final exchange code: mag-shaped layer (T exchange) with the rotated vs the ascending push order, arxiv-shaped layer.
set -x
mkdir -p gpurun_out
T=r02n
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['ms_per_step'],4), (d.get('parity_check') or {}).get('ok'), d.get('single_gpu_same_graph'), 'e2e', (d.get('e2e') or {}).get('ms_per_step'))
    print({k: round(v['ms_per_step'],4) for k,v in d.get('kernels_rank0',{}).items()})
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
}
for ord in rotated ascending; do
  EGC_PEER_ORDER=$ord timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 8 --workload mag --steps 20 --warmup 5 --no-extras > gpurun_out/${T}_bench_8gpu_mag_$ord.json 2> gpurun_out/${T}_bench_8gpu_mag_$ord.err
  echo "mag order=$ord rc=$?"
  show gpurun_out/${T}_bench_8gpu_mag_$ord.json
done
for ord in rotated ascending; do
EGC_PEER_ORDER=$ord timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 \
  bench.py --gpus 8 --steps 20 --warmup 5 --no-extras > gpurun_out/${T}_bench_8gpu_arxiv_$ord.json 2> gpurun_out/${T}_bench_8gpu_arxiv_$ord.err
echo "arxiv rc=$?"
show gpurun_out/${T}_bench_8gpu_arxiv_$ord.json
done
tail -c 600 gpurun_out/${T}_bench_8gpu_mag_rotated.err
