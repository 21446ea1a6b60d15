#!/bin/bash
# Checkpoint r02m on 2 GPUs: new tests (agg_init, overlapped exchanges), mag partitioned over 2 GPUs with the exchanges
# overlapped (copy engine + split launches) vs not, rotated vs ascending push order on the arxiv layer.
set -x
mkdir -p gpurun_out
T=r02m
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_round2.py tests/test_epilogue.py -m gpu -q -rs -x > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -12 gpurun_out/${T}_pytest.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['ms_per_step'],4), (d.get('parity_check') or {}).get('ok'), (d.get('parity_check') or {}).get('max_rel_err'), d.get('single_gpu_same_graph'), 'e2e', (d.get('e2e') or {}).get('ms_per_step'))
    print({k: round(v['ms_per_step'],4) for k,v in d.get('kernels_rank0',{}).items()})
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
}
for ov in 1 0; do
  EGC_DIST_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 2 --workload mag --steps 20 --warmup 5 --no-extras > gpurun_out/${T}_bench_2gpu_mag_ov$ov.json 2> gpurun_out/${T}_bench_2gpu_mag_ov$ov.err
  echo "mag overlap=$ov rc=$?"
  show gpurun_out/${T}_bench_2gpu_mag_ov$ov.json
done
for ord in rotated ascending; do
  EGC_PEER_ORDER=$ord timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > gpurun_out/${T}_bench_2gpu_arxiv_$ord.json 2> gpurun_out/${T}_bench_2gpu_arxiv_$ord.err
  echo "arxiv order=$ord rc=$?"
  show gpurun_out/${T}_bench_2gpu_arxiv_$ord.json
done
tail -c 800 gpurun_out/${T}_bench_2gpu_mag_ov1.err
