#!/bin/bash
# parity tests, then host-side timing of the query phases for c5 / c3 (1B rows) / c2
mkdir -p gpurun_out
T=${TAG:-r2o}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in c5 c3 c2; do
  R=100000000; [ $w = c3 ] && R=1000000000
  SG_HOST_TIMING=1 timeout 600 python bench.py --workload $w --rows $R --steps 6 --warmup 3 --no-e2e --no-cpu --extra none \
    > gpurun_out/${T}_$w.json 2> gpurun_out/${T}_$w.err
  grep "sg host" gpurun_out/${T}_$w.err | tail -2
  python -c "
import json; d=json.load(open('gpurun_out/${T}_$w.json')); print('$w', d['value']/1e9, 'ms/step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms_per_launch'], d['parity']['ok'])"
done
