#!/bin/bash
# C3 / C4 aggregation-pass variants on the current binary (200M rows)
mkdir -p gpurun_out
T=${TAG:-r2r}
run() {  # name, workload, env...
  name=$1; w=$2; shift; shift
  env "$@" SG_PHASE_TIMING=1 timeout 300 python bench.py --workload $w --rows 200000000 --steps 5 --warmup 3 --no-e2e --no-cpu --extra none \
     > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err
  echo "=== $name"; grep -E "sg phase|sg pass" gpurun_out/${T}_$name.err | tail -2
  python -c "
import json
d=json.load(open('gpurun_out/${T}_$name.json'))
print('   kernel ms', d['roofline']['kernel_ms_per_launch'], 'frac', d['roofline']['frac'], 'ms/step', d['ms_per_step'], 'parity', d['parity'] and d['parity']['ok'])" || tail -3 gpurun_out/${T}_$name.err
}
run c3 c3 X=1
run c3_avg c3 SG_BENCH_OP=avg
run c3_onegroup c3 SG_BENCH_GROUPS=d
run c3_nogroup c3 SG_BENCH_GROUPS=
run c4 c4 X=1
run c2 c2 X=1
