#!/bin/bash
# final single-GPU check of the round: parity tests, smoke, the default bench line (C3 at 1B rows, with e2e /
# cpu_baseline / extras) and the reference arm
mkdir -p gpurun_out
T=${TAG:-r02}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${T}_pytest_final.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
tail -3 gpurun_out/${T}_bench_n1.err
python - <<PY
import json
d = json.load(open("gpurun_out/${T}_bench_n1.json"))
print("value %.2f Grows/s ms/step %.3f kernel %.3f frac %.3f traffic %s" % (d["value"] / 1e9, d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"], d["roofline"]["traffic"]))
print("parity", d["parity"]); print("e2e", d["e2e"]); print("cpu", d["cpu_baseline"]); print("clocks", d["clocks"], "launches", d["gpu_launches"])
for x in d.get("extra", []):
    print(x.get("config", {}).get("workload", "")[:30], "value %.2f ms/step %.3f kernel %.3f frac %.3f" % (x.get("value", 0) / 1e9, x.get("ms_per_step", 0), x.get("roofline", {}).get("kernel_ms_per_launch", 0), x.get("roofline", {}).get("frac", 0)), x.get("parity", {}).get("ok"), x.get("error"))
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
cut -c1-400 gpurun_out/${T}_bench_ref.json
