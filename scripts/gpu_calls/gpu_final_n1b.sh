#!/bin/bash
set -u
out=gpurun_out/r2final2
mkdir -p "$out"
export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > "$out/bench_reference_c3.json" 2> "$out/bench_reference_c3.err"; echo "reference exit $?"
timeout 600 python bench.py --steps 20 --warmup 3 > "$out/bench_c3_n1.json" 2> "$out/bench_c3_n1.err"; echo "bench c3 exit $?"
for c in c2 c4 c5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 3 --no-cpu > "$out/bench_${c}_n1.json" 2> "$out/bench_${c}_n1.err"; echo "bench $c exit $?"
done
python - <<P
import json
for c in ("c3","c2","c4","c5"):
    try:
        d=json.loads([l for l in open("$out/bench_%s_n1.json"%c).read().strip().splitlines() if l.startswith("{")][-1])
        r=d["roofline"]; print(c, round(d["value"]), "MLUPS", round(d["ms_per_step"],3), "ms/step frac", round(r["frac"],3), {k:round(v,3) for k,v in r["frac_per_stage"].items()}, "parity", d["parity"] and (d["parity"].get("ok"), d["parity"].get("max_rel")), "e2e", d["e2e"] and (round(d["e2e"]["value"]), round(d["e2e"]["value_loop_only"])), "dropin", d["e2e_dropin"] and round(d["e2e_dropin"]["value"]), "cpu", d["cpu_baseline"] and (d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"]), "launches", d["gpu_launches"], d["clocks"])
        if "whole_stage" in r: print("   whole stage", r["whole_stage"])
    except Exception as e:
        print(c, "parse failed", e)
try:
    d=json.loads(open("$out/bench_reference_c3.json").read().strip().splitlines()[-1]); print("reference", d["value"], d["steps"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["seconds_min_max"])
except Exception as e: print("reference parse failed", e)
P
