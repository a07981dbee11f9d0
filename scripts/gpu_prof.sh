#!/bin/bash
# ncu captures of the scan kernel (and hist_apply) for one workload: W=c3 ROWS=50000000 TAG=r2g
mkdir -p gpurun_out
W=${W:-c3}; ROWS=${ROWS:-50000000}; T=${TAG:-r2g}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scan_kernel|hist_apply" -s 2 -c 2 -f -o gpurun_out/${T}_prof_$W \
  python bench.py --workload $W --rows $ROWS --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity --extra none > gpurun_out/${T}_prof_$W.log 2>&1
tail -3 gpurun_out/${T}_prof_$W.log
ls -la gpurun_out/${T}_prof_$W.ncu-rep
