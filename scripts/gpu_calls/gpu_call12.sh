#!/bin/bash
set -u
out=gpurun_out/r2c12
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -q -m gpu -x > "$out/pytest_gpu.log" 2>&1
echo "pytest exit $?"; tail -6 "$out/pytest_gpu.log"
timeout 600 python bench.py --config c4 --steps 20 --warmup 3 --no-cpu > "$out/bench_c4_n1.json" 2> "$out/bench_c4_n1.err"; echo "bench c4 exit $?"
python - <<P
import json
try:
    d=json.loads([l for l in open("$out/bench_c4_n1.json").read().strip().splitlines() if l.startswith("{")][-1])
    r=d["roofline"]; print("c4", round(d["value"]), "MLUPS", round(d["ms_per_step"],3), "ms/step frac", round(r["frac"],3), {k:round(v,3) for k,v in r["frac_per_stage"].items()}, "parity", d["parity"] and (d["parity"].get("ok"), d["parity"].get("max_rel"), d["parity"].get("error")), "e2e", d["e2e"] and round(d["e2e"]["value"]), d["clocks"])
    print("   whole stage", r["whole_stage"], r["scalar_stage_ms_per_launch"])
except Exception as e:
    print("parse failed", e)
P
tail -3 "$out/bench_c4_n1.err"
