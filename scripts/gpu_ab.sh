#!/bin/bash
# A/B of environment switches on one workload: W=c3 ROWS=200000000 MODES="SG_X=1 ..." (first run = defaults)
mkdir -p gpurun_out
W=${W:-c3}; ROWS=${ROWS:-200000000}; T=${TAG:-ab}
i=0
for mode in "" $MODES; do
  echo "=== $W mode: '$mode'"
  env $mode SG_PHASE_TIMING=1 timeout 300 python bench.py --workload $W --rows $ROWS --steps 5 --warmup 3 --no-e2e --no-cpu --extra none \
     > gpurun_out/${T}_$i.json 2> gpurun_out/${T}_$i.err
  grep -E "sg phase|sg pass" gpurun_out/${T}_$i.err | tail -2
  python -c "
import json
d=json.load(open('gpurun_out/${T}_$i.json'))
print('   kernel ms', d['roofline']['kernel_ms_per_launch'], 'frac', d['roofline']['frac'], 'ms/step', d['ms_per_step'], 'parity', d['parity']['ok'])
"
  i=$((i+1))
done
