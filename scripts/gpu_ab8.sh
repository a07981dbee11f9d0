#!/bin/bash
# 16-warp CTA (1 per SM) against the 8-warp build (2 CTAs per SM): kernel + phase timing
mkdir -p gpurun_out
T=${TAG:-r2s}
for w in ${WL:-c3 c2 c4}; do
  R=200000000; [ $w = c2 ] && R=100000000
  for lib in 16 8; do
    L=sybil_b200/csrc/libsybilgpu.so; [ $lib = 8 ] && L=sybil_b200/csrc/libsybilgpu8.so
    env SG_LIB=$PWD/$L SG_PHASE_TIMING=1 timeout 600 python bench.py --workload $w --rows $R --steps 5 --warmup 3 --no-e2e --no-cpu --extra none \
      > gpurun_out/${T}_${w}_$lib.json 2> gpurun_out/${T}_${w}_$lib.err
    echo "=== $w warps/CTA $lib"; grep -E "sg phase|sg pass" gpurun_out/${T}_${w}_$lib.err | tail -2
    python -c "
import json; d=json.load(open('gpurun_out/${T}_${w}_$lib.json')); print('  ', d['value']/1e9, 'Grows/s ms/step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms_per_launch'], 'frac', d['roofline']['frac'], d['parity'] and d['parity']['ok'])" || tail -5 gpurun_out/${T}_${w}_$lib.err
  done
done
