#!/bin/bash
# A/B runs behind the environment switches (results: DESIGN.md "Tried and dropped" / section 4, profiles/r02m - r02r).
#   bash tools/gpu_ab.sh hints        1 GPU : EGC_BWD_NEAR_MB = 0 32 64 128, arxiv / mag / uniform graph        (r02o)
#   bash tools/gpu_ab.sh ring         1 GPU : EGC_BWD_RING = 0 4 7, arxiv with parity_check, uniform graph     (r02r)
#   bash tools/gpu_ab.sh overlap N    N GPUs: EGC_DIST_OVERLAP = 1 0 on the mag-shaped layer                   (r02m, N = 2)
#   bash tools/gpu_ab.sh order N      N GPUs: EGC_PEER_ORDER = rotated ascending, mag- and arxiv-shaped layers (r02n, N = 8)
set -x
mkdir -p gpurun_out
X=${1:-ring}
N=${2:-2}
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d.get('kernels') or d.get('kernels_rank0') or {}
    print(sys.argv[1], 'ms', round(d['ms_per_step'],4), 'parity', (d.get('parity_check') or {}).get('ok'), d.get('single_gpu_same_graph'))
    print({n: round(v['ms_per_step'],4) for n,v in k.items()})
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
}
one() { env "$1" timeout 300 python bench.py ${@:3} --steps 20 --warmup 5 --no-extras > gpurun_out/ab_$2.json 2> gpurun_out/ab_$2.err; show gpurun_out/ab_$2.json; }
many() { env "$1" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus $N ${@:3} --steps 20 --warmup 5 --no-extras > gpurun_out/ab_$2.json 2> gpurun_out/ab_$2.err; show gpurun_out/ab_$2.json; }
case $X in
  hints)
    for mb in 0 32 64 128; do for w in arxiv mag; do one EGC_BWD_NEAR_MB=$mb near${mb}_$w --workload $w --no-cpu-baseline; done; done
    for mb in 0 64; do one EGC_BWD_NEAR_MB=$mb near${mb}_uniform --locality 0 --no-cpu-baseline; done ;;
  ring)
    EGC_BWD_RING=4 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_epilogue.py -m gpu -q -x | tail -3
    for r in 0 4 7; do one EGC_BWD_RING=$r ring$r; done
    for r in 0 4; do one EGC_BWD_RING=$r ring${r}_uniform --locality 0 --no-cpu-baseline; done ;;
  overlap)
    for ov in 1 0; do many EGC_DIST_OVERLAP=$ov overlap${ov}_mag --workload mag; done ;;
  order)
    for o in rotated ascending; do many EGC_PEER_ORDER=$o order_${o}_mag --workload mag; many EGC_PEER_ORDER=$o order_${o}_arxiv; done ;;
esac
