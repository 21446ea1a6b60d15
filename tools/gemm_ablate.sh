#!/bin/bash
for ab in 0 1 2 4 8 16 3 7 15 14; do
  echo "== EGC_TC_ABLATE=$ab"; EGC_TC_ABLATE=$ab timeout 120 python tools/exp_gemm.py arxiv 5 2>&1 | grep 3xtf32
done
