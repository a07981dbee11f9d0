#!/bin/bash
# parity tests first (stop on a failure), then the ncu evidence of round 2
mkdir -p gpurun_out
T=${TAG:-r02}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${T}_pytest.log
grep -q " passed" gpurun_out/${T}_pytest.log && ! grep -q "failed" gpurun_out/${T}_pytest.log || exit 1
bash scripts/gpu_prof_final.sh
