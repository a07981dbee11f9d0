#!/bin/bash
# two GPUs: selfcheck (merged result vs oracle, both dictionary modes), C3 1B strong-scaled, C5 (large-plan merge)
mkdir -p gpurun_out
T=${TAG:-r02n2}
N=${N:-2}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 8 --warmup 3 --e2e-steps 1 --no-cpu --extra ${EXTRA:-c5,c2} > gpurun_out/${T}.json 2> gpurun_out/${T}.err
tail -5 gpurun_out/${T}.err
python - <<PY
import json
d = json.load(open("gpurun_out/${T}.json"))
print("value %.2f Grows/s ms/step %.3f kernel %.3f frac %.3f parity %s selfcheck %s e2e %s" % (
    d["value"] / 1e9, d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"], d["parity"], d["multi_gpu_selfcheck"], d.get("e2e")))
for x in d.get("extra", []):
    print(x.get("config", {}).get("workload"), x.get("value"), x.get("ms_per_step"), x.get("parity"), x.get("error"))
PY
