#!/bin/bash
mkdir -p gpurun_out
for mode in "" "SG_NO_PUSHDOWN=1" "SG_NO_FAIL_MODE=1" "SG_NO_HIST_CACHE=1"; do
  echo "=== mode: $mode"
  env $mode timeout 200 python scripts/diag_c3.py c3 400000 2>&1 | tail -12
done > gpurun_out/r2c_diag.txt 2>&1
cat gpurun_out/r2c_diag.txt
