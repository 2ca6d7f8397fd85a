#!/bin/bash
set -u
out=gpurun_out/r2c8
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_api.py tests/test_gpu_kernels.py tests/test_gpu_config_parity.py -q -m gpu -x 2>&1 | tail -4
timeout 600 python scripts/dev_time_r2.py gwvar > "$out/time_gwvar.log" 2>&1; grep -v "^c4 scalar part" "$out/time_gwvar.log"
