#!/bin/bash
# ncu evidence (1 GPU): launch list of two eager steps + one `--set full` capture of every step kernel, per workload.
# Numbers printed by bench.py under ncu are NOT bench values.      usage: bash tools/gpu_r02_ncu.sh <tag> [workloads]
# (r02f: arxiv mag; r02p: arxiv).  Summarise with tools/ncu_summary.py.
set -x
mkdir -p gpurun_out
T=${1:-r02f}
W=${2:-"arxiv mag"}
RX='regex:k_aggregate_rows|k_scatter_cols|k_scatter_ring|k_combine_bwd|k_route_minmax|k_project_tc|k_wgrad_mn|k_wgrad_tc'
for w in $W; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/${T}_launches_$w.csv \
    python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-graph > gpurun_out/${T}_ncu_launch_$w.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k "$RX" -s 36 -c 9 -f -o gpurun_out/${T}_full_$w \
    python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-graph > gpurun_out/${T}_ncu_full_$w.log 2>&1
  ls -la gpurun_out/${T}_full_$w.ncu-rep
done
head -20 gpurun_out/${T}_launches_arxiv.csv
