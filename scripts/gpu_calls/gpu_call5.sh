#!/bin/bash
set -u
out=gpurun_out/r2c5
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -m gpu -x > "$out/pytest_gpu.log" 2>&1
echo "pytest exit $?"; tail -6 "$out/pytest_gpu.log"
timeout 300 python scripts/dev_time_r2.py chunk > "$out/time_chunk.log" 2>&1; cat "$out/time_chunk.log"
timeout 600 python scripts/dev_time_r2.py gwvar > "$out/time_gwvar.log" 2>&1; grep -v "^c4 scalar part" "$out/time_gwvar.log"
timeout 300 python scripts/dev_time_api.py > "$out/time_api.log" 2>&1; cat "$out/time_api.log"
for c in c4 c2; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 --no-cpu > "$out/bench_$c.json" 2> "$out/bench_$c.err"
  echo "bench $c exit $?"; python - <<P
import json
try:
    d=json.loads(open("$out/bench_$c.json").read().strip().splitlines()[-1])
    r=d["roofline"]; print("$c", round(d["value"]), "MLUPS", round(d["ms_per_step"],3), "ms/step frac", round(r["frac"],3), {k:round(v,3) for k,v in r["frac_per_stage"].items()}, "parity", d["parity"] and (d["parity"].get("ok"), d["parity"].get("max_rel"), d["parity"].get("error")), "e2e", d["e2e"] and round(d["e2e"]["value"]), "dropin", d["e2e_dropin"] and round(d["e2e_dropin"]["value"]), "launches", d["gpu_launches"])
    if "whole_stage" in r: print("   whole stage", r["whole_stage"])
except Exception as e:
    print("$c parse failed", e)
P
  tail -3 "$out/bench_$c.err"
done
