#!/bin/bash
# round-end GPU call (1 GPU): parity suite, smoke, both bench workloads (full contract lines), reference arm,
# ncu launch list + full capture of the step kernels.  Outputs under gpurun_out/ with the given tag.
TAG=${1:-r01e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -2 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_1gpu_arxiv.json 2> gpurun_out/${TAG}_bench_arxiv.err; echo "bench arxiv rc=$?"
timeout 600 python bench.py --workload mag --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_1gpu_mag.json 2> gpurun_out/${TAG}_bench_mag.err; echo "bench mag rc=$?"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference arm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none \
  -k regex:'k_scatter_cols|k_aggregate_rows|k_combine_bwd|k_route_minmax|k_project_tc|k_wgrad_tc|k_wgrad_reduce|k_colsum' -c 10 -f -o gpurun_out/${TAG}_full \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu full rc=$?"
python - <<PY
import json
for f in ("${TAG}_bench_1gpu_arxiv", "${TAG}_bench_1gpu_mag", "${TAG}_bench_reference"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("ms_per_step"), d.get("value"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
