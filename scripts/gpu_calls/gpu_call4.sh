#!/bin/bash
set -u
out=gpurun_out/r2c4
mkdir -p "$out"
export PYTHONUNBUFFERED=1
prof() {  # name, kernel regex, skip, python snippet
  local name=$1 regex=$2 skip=$3 code=$4
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 \
     -o "/tmp/$name" python -c "$code" > "$out/ncu_$name.log" 2>&1
  echo "ncu $name exit $?"
  ncu -i /tmp/$name.ncu-rep --page raw --csv > "$out/${name}_raw.csv" 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page details --csv > "$out/${name}_details.csv" 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv > "$out/${name}_source.csv" 2>/dev/null
  gzip -f "$out/${name}_source.csv"
  ls -la /tmp/$name.ncu-rep
}
PRE='import sys; sys.path.insert(0, "scripts"); sys.argv=["x","none"]; import dev_time_r2 as D, torch; '
prof gw_f32 tensor_stage_ws 4 "$PRE D.gw_case(384, torch.float32)"
prof scalar_f32_nf2 fused_stage_ws 3 "$PRE D.scalar_case((384,)*3, torch.float32, 1, True, nscalars=2)"
prof scalar_f64_stage0 fused_stage_ws 3 "$PRE D.scalar_case((512,)*3, torch.float64, 0, True)"
prof scalar_f64_stage1 fused_stage_ws 3 "$PRE D.scalar_case((512,)*3, torch.float64, 1, True)"
prof scalar_f64_64planes fused_stage_ws 3 "$PRE D.scalar_case((64,512,512), torch.float64, 1, True)"
timeout 600 python scripts/dev_time_r2.py gwvar > "$out/time_gwvar.log" 2>&1
du -sh "$out"
grep -v "^c4 scalar part" "$out/time_gwvar.log"
