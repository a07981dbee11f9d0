#!/bin/bash
# parity tests with the default variant choice and with the 8-warp build forced, then kernel timings per variant
mkdir -p gpurun_out
T=${TAG:-r2t}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${T}_pytest.log
SG_VARIANT=8 timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${T}_pytest_w8.log
for w in ${WL:-c3 c2 c4 c5}; do
  R=200000000; [ $w = c2 ] && R=100000000
  for v in default 8 16; do
    E=X=1; [ $v != default ] && E=SG_VARIANT=$v
    env $E SG_PHASE_TIMING=1 timeout 600 python bench.py --workload $w --rows $R --steps 5 --warmup 3 --no-e2e --no-cpu --extra none \
      > gpurun_out/${T}_${w}_$v.json 2> gpurun_out/${T}_${w}_$v.err
    echo "=== $w variant $v"; grep -E "sg phase|sg pass" gpurun_out/${T}_${w}_$v.err | tail -2
    python -c "
import json; d=json.load(open('gpurun_out/${T}_${w}_$v.json')); print('  ', d['value']/1e9, 'Grows/s ms/step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms_per_launch'], 'frac', d['roofline']['frac'], d['parity'] and d['parity']['ok'])" || tail -5 gpurun_out/${T}_${w}_$v.err
  done
done
