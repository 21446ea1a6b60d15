#!/bin/bash
# Checkpoint r02k on 2 GPUs: whole GPU suite, default 1-GPU bench line, the mag-shaped layer partitioned over 2 GPUs
# (T-exchange backward; EGC_DIST_T_EXCHANGE=0 = the partial-sum exchange for comparison), REGConv bench.
set -x
mkdir -p gpurun_out
T=r02k
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -15 gpurun_out/${T}_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
head -c 300 gpurun_out/${T}_bench_1gpu.json; echo
for mode in auto 0; do
  EGC_DIST_T_EXCHANGE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 2 --workload mag --steps 20 --warmup 5 --no-extras > gpurun_out/${T}_bench_2gpu_mag_t$mode.json 2> gpurun_out/${T}_bench_2gpu_mag_t$mode.err
  echo "mag t=$mode rc=$?"
  python - gpurun_out/${T}_bench_2gpu_mag_t$mode.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['ms_per_step'],4), d.get('parity_check'), d.get('single_gpu_same_graph'))
    print({k: round(v['ms_per_step'],4) for k,v in d.get('kernels_rank0',{}).items()})
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
timeout 400 python bench.py --workload rmag --steps 5 --warmup 3 > gpurun_out/${T}_bench_rmag.json 2> gpurun_out/${T}_bench_rmag.err
head -c 600 gpurun_out/${T}_bench_rmag.json; echo
tail -c 600 gpurun_out/${T}_bench_2gpu_mag_tauto.err
