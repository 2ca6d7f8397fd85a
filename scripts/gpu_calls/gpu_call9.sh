#!/bin/bash
set -u
out=gpurun_out/r2c9
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 600 python scripts/dev_time_r2.py gwvar > "$out/time_gwvar.log" 2>&1; grep -v "^c4 scalar part" "$out/time_gwvar.log"
