#!/bin/bash
set -u
out=gpurun_out/r2c7
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -m gpu -x > "$out/pytest_gpu.log" 2>&1
echo "pytest exit $?"; tail -6 "$out/pytest_gpu.log"
timeout 600 python scripts/dev_time_r2.py gwvar > "$out/time_gwvar.log" 2>&1; grep -v "^c4 scalar part" "$out/time_gwvar.log"
PRE='import sys; sys.path.insert(0, "scripts"); sys.argv=["x","none"]; import dev_time_r2 as D, torch; '
timeout 300 ncu --set full --clock-control none -k regex:tensor_stage_ws -s 4 -c 1 -o /tmp/gw python -c "$PRE D.gw_case(384, torch.float32)" > "$out/ncu_gw.log" 2>&1
ncu -i /tmp/gw.ncu-rep --page details --csv > "$out/gw_f32_details.csv" 2>/dev/null
ncu -i /tmp/gw.ncu-rep --page raw --csv > "$out/gw_f32_raw.csv" 2>/dev/null
