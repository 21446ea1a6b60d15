#!/bin/bash
# Round 2, GPU call D (2 GPUs): fixed-shape mini-batch path + DP arm; mag / EGC-S backward split A/B at 2 GPUs.
set -x
mkdir -p gpurun_out
T=${1:-r02d}
timeout 600 python -m pytest tests/test_batch.py tests/test_gpu_round2.py tests/test_gpu_dist.py -m gpu -q -rs > gpurun_out/${T}_pytest_new.log 2>&1
echo "pytest new rc=$?" >> gpurun_out/${T}_pytest_new.log
for w in zinc cifar; do
  timeout 300 python bench.py --workload $w --steps 24 --warmup 5 > gpurun_out/${T}_bench_1gpu_$w.json 2> gpurun_out/${T}_bench_1gpu_$w.err
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --workload $w --steps 24 --warmup 5 > gpurun_out/${T}_bench_2gpu_$w.json 2> gpurun_out/${T}_bench_2gpu_$w.err
done
run2() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --workload mag --steps 20 --warmup 5 --no-extras > gpurun_out/${T}_2gpu_mag_$name.json 2> gpurun_out/${T}_2gpu_mag_$name.err
  echo "$name rc=$?"
}
run2 nosplit EGC_DIST_SPLIT_BWD=0
run2 split EGC_DIST_SPLIT_BWD=1
run2 split_push1 EGC_DIST_SPLIT_BWD=1 EGC_PEER_PUSH_CTAS_PER_SM=1
tail -4 gpurun_out/${T}_pytest_new.log
for f in gpurun_out/${T}_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['ms_per_step'],4), (d.get('parity_check') or {}).get('ok'), d.get('eager'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
tail -c 600 gpurun_out/${T}_bench_1gpu_zinc.err gpurun_out/${T}_bench_2gpu_zinc.err
