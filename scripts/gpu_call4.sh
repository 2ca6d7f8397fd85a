#!/bin/bash
set -u
out=gpurun_out/r2c4
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 600 python scripts/dev_time_r2.py gwvar 2>&1 | tee "$out/time_gwvar.log"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tensor_stage_ws -s 4 -c 1 -o "$out/gw_f32" python scripts/dev_time_r2.py gwprof > "$out/ncu_gw.log" 2>&1
echo "ncu gw exit $?"; tail -3 "$out/ncu_gw.log"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_stage_ws -s 3 -c 1 -o "$out/scalar_f32_nf2" python - > "$out/ncu_nf2.log" 2>&1 <<P
import sys; sys.path.insert(0, "scripts"); sys.argv=["x","none"]
import dev_time_r2 as D, torch
D.scalar_case((384,)*3, torch.float32, 1, True, nscalars=2)
P
echo "ncu nf2 exit $?"; tail -3 "$out/ncu_nf2.log"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_stage_ws -s 3 -c 1 -o "$out/scalar_f64_stage0" python - > "$out/ncu_s0.log" 2>&1 <<P
import sys; sys.path.insert(0, "scripts"); sys.argv=["x","none"]
import dev_time_r2 as D, torch
D.scalar_case((512,)*3, torch.float64, 0, True)
P
echo "ncu stage0 exit $?"; tail -3 "$out/ncu_s0.log"
ls -la "$out"
