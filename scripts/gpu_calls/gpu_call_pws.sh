#!/bin/bash
set -u
out=gpurun_out/r2pws
mkdir -p "$out"
export PYTHONUNBUFFERED=1
# correctness first (bench parity block + a few steps), each under its own timeout
for c in c2 c5; do
  timeout 180 python bench.py --config $c --variant 30 --steps 4 --warmup 3 --no-e2e --no-cpu > "$out/bench_${c}_pws.json" 2> "$out/bench_${c}_pws.err"
  echo "bench $c pws exit $?"
  python - <<P
import json
try:
    d=json.loads([l for l in open("$out/bench_${c}_pws.json").read().strip().splitlines() if l.startswith("{")][-1])
    r=d["roofline"]; print("$c pws", round(d["value"]), "MLUPS", round(d["ms_per_step"],4), "frac", round(r["frac"],3), {k:round(v,3) for k,v in r["frac_per_stage"].items()}, "parity", d["parity"] and (d["parity"].get("ok"), d["parity"].get("max_rel"), d["parity"].get("error")))
except Exception as e:
    print("$c parse failed", e)
P
  tail -3 "$out/bench_${c}_pws.err"
done
timeout 300 python scripts/dev_time_r2.py pws > "$out/time_pws.log" 2>&1; cat "$out/time_pws.log"
