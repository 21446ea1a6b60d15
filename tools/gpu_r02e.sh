#!/bin/bash
# Round 2, GPU call E (2 GPUs): MN-major wgrad kernel (tests + A/B), piece-wise push / reduce kernels, DP graph path.
set -x
mkdir -p gpurun_out
T=${1:-r02e}
timeout 900 python -m pytest tests -m gpu -q -rs -x > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
EGC_WGRAD_TRANSPOSE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/${T}_bench_1gpu_wgtr.json 2> gpurun_out/${T}_bench_1gpu_wgtr.err
timeout 300 python bench.py --workload mag --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/${T}_bench_1gpu_mag.json 2> gpurun_out/${T}_bench_1gpu_mag.err
for w in zinc cifar; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --workload $w --steps 24 --warmup 5 > gpurun_out/${T}_bench_2gpu_$w.json 2> gpurun_out/${T}_bench_2gpu_$w.err
done
for w in arxiv mag; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --workload $w --steps 20 --warmup 5 --no-extras > gpurun_out/${T}_2gpu_$w.json 2> gpurun_out/${T}_2gpu_$w.err
done
tail -6 gpurun_out/${T}_pytest.log
for f in gpurun_out/${T}_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['ms_per_step'],4), (d.get('parity_check') or {}).get('ok'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
tail -c 800 gpurun_out/${T}_bench_2gpu_zinc.err
