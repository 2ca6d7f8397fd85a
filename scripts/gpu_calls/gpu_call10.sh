#!/bin/bash
set -u
out=gpurun_out/r2c10
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py tests/test_gpu_config_parity.py -q -m gpu -x 2>&1 | tail -4
timeout 300 python scripts/dev_time_r2.py ab 2>&1 | tee "$out/time_ab.log"
