#!/bin/bash
# round-2 ncu evidence: launch list of the default bench command + one full capture of the scan kernel per config.
# The captures are ~40 MB each and gpurun brings back at most 64 MiB: they are summarised on the box
# (scripts/ncu_summary.py) and only the headline config's report is kept.
mkdir -p gpurun_out
T=${TAG:-r02}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${T}_launches_c3.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --extra none > gpurun_out/${T}_launches_c3.log 2>&1
tail -2 gpurun_out/${T}_launches_c3.log | cut -c1-300
for w in ${WL:-c3 c2 c4 c5}; do
  R=200000000; [ $w = c2 ] && R=100000000; [ $w = c5 ] && R=100000000
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -f -o gpurun_out/${T}_prof_$w \
    python bench.py --workload $w --rows $R --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity --extra none > gpurun_out/${T}_prof_$w.log 2>&1
  python scripts/ncu_summary.py gpurun_out/${T}_prof_$w.ncu-rep > gpurun_out/${T}_summary_$w.md 2> gpurun_out/${T}_summary_$w.err
  head -30 gpurun_out/${T}_summary_$w.md
  [ $w != ${KEEP:-c3} ] && rm -f gpurun_out/${T}_prof_$w.ncu-rep
done
ls -la gpurun_out
