#!/bin/bash
# parity tests, then kernel + phase timing for c2 / c3 / c4 / c5 at 200M rows (narrow arrays unless SG_WIDE=1)
mkdir -p gpurun_out
T=${TAG:-r2q}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${T}_pytest.log
for w in ${WL:-c3 c2 c4 c5}; do
  R=200000000; [ $w = c2 ] && R=100000000
  for mode in narrow wide; do
    E=X=1; [ $mode = wide ] && E=SG_WIDE=1
    env $E SG_PHASE_TIMING=1 timeout 600 python bench.py --workload $w --rows $R --steps 5 --warmup 3 --no-e2e --no-cpu --extra none \
      > gpurun_out/${T}_${w}_$mode.json 2> gpurun_out/${T}_${w}_$mode.err
    echo "=== $w $mode"; grep -E "sg phase|sg pass" gpurun_out/${T}_${w}_$mode.err | tail -2
    python -c "
import json; d=json.load(open('gpurun_out/${T}_${w}_$mode.json')); print('  ', d['value']/1e9, 'Grows/s ms/step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms_per_launch'], 'frac', d['roofline']['frac'], 'resident', d['roofline']['encoded_bytes_resident'], d['parity'] and d['parity']['ok'])" || tail -5 gpurun_out/${T}_${w}_$mode.err
  done
done
