#!/bin/bash
set -u
out=gpurun_out/r2c11
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 400 python scripts/dev_time_r2.py jit2 2>&1 | tee "$out/time_jit2.log"
