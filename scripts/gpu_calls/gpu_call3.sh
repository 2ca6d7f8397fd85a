#!/bin/bash
set -u
out=gpurun_out/r2c3
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -m gpu -x > "$out/pytest_gpu.log" 2>&1
echo "pytest exit $?"; tail -15 "$out/pytest_gpu.log"
echo "== A: default lib"; timeout 300 python scripts/dev_time_r2.py ab 2>&1 | tee "$out/time_ab_default.log"
echo "== B: per-site poly"; PSB_DEV_LIB=libpsb_polyold.so timeout 300 python scripts/dev_time_r2.py ab 2>&1 | tee "$out/time_ab_polyold.log"
timeout 300 python scripts/dev_time_r2.py chunk 2>&1 | tee "$out/time_chunk.log"
for c in c3 c2 c4 c5; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 --no-cpu > "$out/bench_$c.json" 2> "$out/bench_$c.err"
  echo "bench $c exit $?"; python - <<P
import json
try:
    d=json.loads(open("$out/bench_$c.json").read().strip().splitlines()[-1])
    r=d["roofline"]; print("$c", round(d["value"]), "MLUPS", round(d["ms_per_step"],3), "ms/step frac", round(r["frac"],3), {k:round(v,3) for k,v in r["frac_per_stage"].items()}, "parity", d["parity"] and (d["parity"].get("ok"), d["parity"].get("max_rel"), d["parity"].get("error")), "e2e", d["e2e"] and round(d["e2e"]["value"]), "launches", d["gpu_launches"])
    if "whole_stage" in r: print("   whole stage", r["whole_stage"])
except Exception as e:
    print("$c parse failed", e)
P
  tail -3 "$out/bench_$c.err"
done
