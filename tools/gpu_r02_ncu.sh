#!/bin/bash
# Round 2 ncu evidence (1 GPU): launch list of two eager steps + one `--set full` capture of every step kernel, for the
# arxiv-shaped (default) and mag-shaped workloads.  Numbers printed by bench.py under ncu are NOT bench values.
set -x
mkdir -p gpurun_out
T=${1:-r02f}
RX='regex:k_aggregate_rows|k_scatter_cols|k_combine_bwd|k_route_minmax|k_project_tc|k_wgrad_mn|k_wgrad_tc'
for w in arxiv mag; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/${T}_launches_$w.csv \
    python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-graph > gpurun_out/${T}_ncu_launch_$w.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k "$RX" -s 36 -c 9 -f -o gpurun_out/${T}_full_$w \
    python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-graph > gpurun_out/${T}_ncu_full_$w.log 2>&1
  ls -la gpurun_out/${T}_full_$w.ncu-rep
done
tail -3 gpurun_out/${T}_ncu_full_arxiv.log
head -20 gpurun_out/${T}_launches_arxiv.csv
