#!/bin/bash
# one GPU call: full GPU suite + ncu full capture (with source) of the row-block forward and column-block backward kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/c2_pytest.log
tail -3 gpurun_out/c2_pytest.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_scatter_cols|k_aggregate_rows|k_combine_bwd|k_route_minmax' -c 4 -f -o gpurun_out/c2_full \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c2_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | tail -5
