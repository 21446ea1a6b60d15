#!/bin/bash
# Scaling record on N GPUs of one box: the default partitioned bench line (arxiv 1 layer + other_configs: 3-layer stack, mag)
# and the data-parallel mini-batch arms.   usage: bash tools/gpu_scale.sh <tag> <N>
set -x
mkdir -p gpurun_out
T=${1:-r02g}
N=${2:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err
echo "bench rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus $N --workload mag --steps 20 --warmup 5 --no-extras > gpurun_out/${T}_bench_${N}gpu_mag.json 2> gpurun_out/${T}_bench_${N}gpu_mag.err
echo "mag rc=$?"
for w in cifar zinc; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --workload $w --steps 24 --warmup 5 > gpurun_out/${T}_bench_${N}gpu_$w.json 2> gpurun_out/${T}_bench_${N}gpu_$w.err
  echo "$w rc=$?"
done
for f in gpurun_out/${T}_bench_${N}gpu*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['ms_per_step'],4), (d.get('parity_check') or {}).get('ok'), d.get('single_gpu_same_graph'))
    for k,v in (d.get('other_configs') or {}).items(): print('   ', k, v.get('ms_per_step'), v.get('single_gpu_same_graph'), (v.get('parity_check') or {}).get('ok'), v.get('error'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
tail -c 1500 gpurun_out/${T}_bench_${N}gpu.err
