#!/bin/bash
# gpurun --gpus 2 -- 'bash scripts/gpu_calls/gpu_call_n2.sh'
set -u
out=gpurun_out/r2n2
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_distributed.py -q -m gpu -x > "$out/pytest_dist.log" 2>&1
echo "pytest dist exit $?"; tail -4 "$out/pytest_dist.log"
