#!/bin/bash
# two GPUs: tests/multi_gpu_check.py (4 queries x seeded / per-rank dictionaries against the oracle)
mkdir -p gpurun_out
T=${TAG:-r02n2c}
N=${N:-2}
timeout 600 env ${ENVX:-X=1} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
  tests/multi_gpu_check.py > gpurun_out/${T}.log 2>&1
grep -E "query|MULTI_GPU|Error|error" gpurun_out/${T}.log | head -30
