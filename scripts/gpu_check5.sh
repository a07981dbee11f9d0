#!/bin/bash
# parity tests, then C5 against the previous commit's library on one box, and the other configs as a sanity line
mkdir -p gpurun_out
T=${TAG:-r2x}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${T}_pytest.log
for w in c5 c5 c3 c2 c4; do
  R=200000000; [ $w = c2 ] && R=100000000
  for mode in prev cur; do
    [ $w != c5 ] && [ $mode = prev ] && continue
    L=sybil_b200/csrc/libsybilgpu.so
    [ $mode = prev ] && L=sybil_b200/csrc/libsybilgpu_prev.so
    env SG_LIB=$PWD/$L SG_PHASE_TIMING=1 timeout 600 python bench.py --workload $w --rows $R --steps 5 --warmup 3 --no-e2e --no-cpu --extra none \
      > gpurun_out/${T}_${w}_$mode.json 2> gpurun_out/${T}_${w}_$mode.err
    echo "=== $w $mode"; grep -E "sg phase|sg pass" gpurun_out/${T}_${w}_$mode.err | tail -2
    python -c "
import json; d=json.load(open('gpurun_out/${T}_${w}_$mode.json')); print('  ', 'kernel', d['roofline']['kernel_ms_per_launch'], 'frac', d['roofline']['frac'], 'ms/step', d['ms_per_step'], d['parity'] and d['parity']['ok'])" || tail -5 gpurun_out/${T}_${w}_$mode.err
  done
done
