#!/bin/bash
# parity tests + phase-timed runs
set -x
mkdir -p gpurun_out
T=${TAG:-r2d}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1
tail -15 gpurun_out/${T}_pytest.log
for w in ${WL:-c3 c4 c2}; do
  SG_PHASE_TIMING=1 timeout 300 python bench.py --workload $w --rows 200000000 --steps 3 --warmup 3 --no-e2e --no-cpu \
    > gpurun_out/${T}_ph_$w.json 2> gpurun_out/${T}_ph_$w.err
  grep "sg phase" gpurun_out/${T}_ph_$w.err | tail -1
  python -c "
import json
d=json.load(open('gpurun_out/${T}_ph_$w.json'))
print('$w', 'G rows/s', d['value']/1e9, 'ms/step', d['ms_per_step'], 'kernel ms', d['roofline']['kernel_ms_per_launch'], 'frac', d['roofline']['frac'], d['parity'])
"
done
free -g | head -2
[ "${FULL:-1}" = "1" ] || exit 0
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -5 gpurun_out/${T}_bench.err
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print(json.dumps({k: d[k] for k in d if k not in ('extra',)}, indent=1)[:3000])
for x in d.get('extra', []): print(json.dumps(x)[:1200])
"
