#!/bin/bash
# Final single-GPU evidence of round 2 (r02p): smoke(), default bench line, fused-epilogue bench, ncu launch list and
# one `--set full` capture per step kernel of the arxiv-shaped step.  Numbers printed under ncu are NOT bench values.
set -x
mkdir -p gpurun_out
T=r02p
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
head -c 300 gpurun_out/${T}_bench_1gpu.json; echo
timeout 200 python tools/epilogue_bench.py > gpurun_out/${T}_epilogue_bench.json 2> gpurun_out/${T}_epilogue_bench.err
cat gpurun_out/${T}_epilogue_bench.json; tail -3 gpurun_out/${T}_epilogue_bench.err
RX='regex:k_aggregate_rows|k_scatter_cols|k_combine_bwd|k_route_minmax|k_project_tc|k_wgrad_mn|k_wgrad_tc'
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/${T}_launches_arxiv.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-graph > gpurun_out/${T}_ncu_launch_arxiv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "$RX" -s 36 -c 9 -f -o gpurun_out/${T}_full_arxiv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-graph > gpurun_out/${T}_ncu_full_arxiv.log 2>&1
ls -la gpurun_out/${T}_full_arxiv.ncu-rep
head -12 gpurun_out/${T}_launches_arxiv.csv
