#!/bin/bash
# where the C3 aggregation pass spends its time: A/B over query variants and cache switches, then an ncu capture
mkdir -p gpurun_out
T=${TAG:-r2p}
run() {  # name, env...
  name=$1; shift
  env "$@" SG_PHASE_TIMING=1 timeout 300 python bench.py --workload ${W:-c3} --rows 200000000 --steps 5 --warmup 3 --no-e2e --no-cpu --no-parity --extra none \
     > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err
  echo "=== $name"; grep -E "sg phase|sg pass" gpurun_out/${T}_$name.err | tail -2
  python -c "
import json
d=json.load(open('gpurun_out/${T}_$name.json'))
print('   kernel ms', d['roofline']['kernel_ms_per_launch'], 'frac', d['roofline']['frac'])"
}
run base X=1
run avg SG_BENCH_OP=avg
run onegroup SG_BENCH_GROUPS=d
run nogroup SG_BENCH_GROUPS=
run nocache SG_NO_HIST_CACHE=1
W=c3 ROWS=50000000 TAG=$T bash scripts/gpu_prof.sh
