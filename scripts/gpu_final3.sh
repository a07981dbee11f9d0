#!/bin/bash
# parity tests + the C3 line with e2e (staging-path check)
mkdir -p gpurun_out
T=${TAG:-r02b}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/${T}_pytest.log
timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu --extra none > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${T}_bench.json"))
print("value %.2f Grows/s ms/step %.3f kernel %.3f frac %.3f" % (d["value"] / 1e9, d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"]))
print("parity", d["parity"]["ok"]); print("e2e", d["e2e"]); print("setup", d["setup"])
PY
