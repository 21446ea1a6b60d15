#!/bin/bash
# one GPU call: parity suite, slab-layout A/B on both bench workloads, GEMM ring-depth diagnostics, ncu capture
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/c1_pytest.log
tail -3 gpurun_out/c1_pytest.log
for f in 64 16 32; do
  timeout 300 python bench.py --no-cpu-baseline --bwd-flags $f > gpurun_out/c1_arxiv_f$f.json 2> gpurun_out/c1_arxiv_f$f.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/c1_arxiv_f$f.json").read().strip().splitlines()[-1])
    print("arxiv flags=$f", d["ms_per_step"], {k: round(v["ms_per_step"], 4) for k, v in d["kernels"].items()})
except Exception as e:
    print("arxiv flags=$f failed", e)
PY
done
for f in 64 16 32; do
  timeout 400 python bench.py --workload mag --no-cpu-baseline --bwd-flags $f > gpurun_out/c1_mag_f$f.json 2> gpurun_out/c1_mag_f$f.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/c1_mag_f$f.json").read().strip().splitlines()[-1])
    print("mag flags=$f", d["ms_per_step"], {k: round(v["ms_per_step"], 4) for k, v in d["kernels"].items()})
except Exception as e:
    print("mag flags=$f failed", e)
PY
done
timeout 200 python tools/exp_gemm.py arxiv 2>&1 | tee gpurun_out/c1_gemm.log
EGC_TC_MAX_RAW_STAGES=6 timeout 200 python tools/exp_gemm.py arxiv 2>&1 | tee -a gpurun_out/c1_gemm.log
EGC_TC_MAX_RAW_STAGES=4 timeout 200 python tools/exp_gemm.py arxiv 2>&1 | tee -a gpurun_out/c1_gemm.log
timeout 200 python tools/exp_gemm.py mag 2>&1 | tee -a gpurun_out/c1_gemm.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_scatter_slab|k_project_tc|k_aggregate_fast' -c 4 -f -o gpurun_out/c1_full \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c1_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out
