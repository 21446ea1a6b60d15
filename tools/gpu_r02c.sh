#!/bin/bash
# Round 2, GPU call C (2 GPUs): new tests; 1-GPU bench with the swizzled GEMM feed (+ A/B without); 2-GPU A/B matrix of the
# backward split / fused signal / push grid.
set -x
mkdir -p gpurun_out
T=${1:-r02c}
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_stack.py tests/test_gpu_fullsize.py tests/test_gpu_dist.py -m gpu -q -rs > gpurun_out/${T}_pytest_new.log 2>&1
echo "pytest new rc=$?" >> gpurun_out/${T}_pytest_new.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
EGC_TC_NO_SWIZZLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/${T}_bench_1gpu_noswz.json 2> gpurun_out/${T}_bench_1gpu_noswz.err
run2() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > gpurun_out/${T}_2gpu_$name.json 2> gpurun_out/${T}_2gpu_$name.err
  echo "$name rc=$?"
}
run2 v0 EGC_DIST_SPLIT_BWD=0 EGC_PEER_FUSED_SIGNAL=0 EGC_PEER_PUSH_CTAS_PER_SM=8
run2 v1 EGC_DIST_SPLIT_BWD=1 EGC_PEER_FUSED_SIGNAL=0 EGC_PEER_PUSH_CTAS_PER_SM=1
run2 v2 EGC_DIST_SPLIT_BWD=1 EGC_PEER_FUSED_SIGNAL=0 EGC_PEER_PUSH_CTAS_PER_SM=8
run2 v3 EGC_DIST_SPLIT_BWD=0 EGC_PEER_FUSED_SIGNAL=1 EGC_PEER_PUSH_CTAS_PER_SM=8
run2 v4 EGC_DIST_SPLIT_BWD=0 EGC_PEER_FUSED_SIGNAL=0 EGC_PEER_PUSH_CTAS_PER_SM=1
tail -4 gpurun_out/${T}_pytest_new.log
for f in gpurun_out/${T}_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['ms_per_step'],4), d.get('parity_check',{}) and d['parity_check'].get('ok'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
