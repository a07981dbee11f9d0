#!/bin/bash
# round-2 "before" measurements: phase cycles for C3/C4/C5, the 1B-row C3 run, ncu captures
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
timeout 120 ./profiles/microbench/hist_red > gpurun_out/r2a_hist_red.txt 2>&1
for w in c3 c4; do
  SG_PHASE_TIMING=1 timeout 300 python bench.py --workload $w --rows 200000000 --steps 3 --warmup 3 --no-e2e --no-cpu \
    > gpurun_out/r2a_ph_$w.json 2> gpurun_out/r2a_ph_$w.err
done
SG_PHASE_TIMING=1 timeout 400 python bench.py --workload c5 --rows 100000000 --steps 2 --warmup 1 --no-e2e --no-cpu \
    > gpurun_out/r2a_ph_c5.json 2> gpurun_out/r2a_ph_c5.err
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2a_c3_1b.json 2> gpurun_out/r2a_c3_1b.err
for w in c3 c4; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -f -o gpurun_out/r2a_prof_$w \
    python bench.py --workload $w --rows 50000000 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2a_prof_$w.log 2>&1
done
ls -la gpurun_out
