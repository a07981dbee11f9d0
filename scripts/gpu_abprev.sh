#!/bin/bash
# regression check on ONE box: the previous commit's library against the current one (16-warp build forced)
mkdir -p gpurun_out
T=${TAG:-r2u}
for w in ${WL:-c2 c3 c5}; do
  R=200000000; [ $w = c2 ] && R=100000000
  for lib in prev cur prev cur; do
    L=sybil_b200/csrc/libsybilgpu.so; [ $lib = prev ] && L=sybil_b200/csrc/libsybilgpu_prev.so
    env SG_LIB=$PWD/$L SG_VARIANT=16 SG_PHASE_TIMING=1 timeout 600 python bench.py --workload $w --rows $R --steps 5 --warmup 3 --no-e2e --no-cpu --extra none \
      > gpurun_out/${T}_${w}_$lib.json 2> gpurun_out/${T}_${w}_$lib.err
    echo "=== $w $lib"; grep -E "sg phase|sg pass" gpurun_out/${T}_${w}_$lib.err | tail -2
    python -c "
import json; d=json.load(open('gpurun_out/${T}_${w}_$lib.json')); print('  ', 'kernel', d['roofline']['kernel_ms_per_launch'], 'frac', d['roofline']['frac'], d['parity'] and d['parity']['ok'])" || tail -5 gpurun_out/${T}_${w}_$lib.err
  done
done
