#!/bin/bash
set -u
out=gpurun_out/r2c13
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -q -m gpu -x > "$out/pytest_gpu.log" 2>&1
echo "pytest exit $?"; tail -5 "$out/pytest_gpu.log"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 > "$out/bench_c3_n1.json" 2> "$out/bench_c3_n1.err"; echo "bench c3 exit $?"
timeout 600 python bench.py --config c5 --steps 20 --warmup 3 --no-cpu > "$out/bench_c5_n1.json" 2> "$out/bench_c5_n1.err"; echo "bench c5 exit $?"
python - <<P
import json
for c in ("c3","c5"):
    try:
        d=json.loads([l for l in open("$out/bench_%s_n1.json"%c).read().strip().splitlines() if l.startswith("{")][-1])
        r=d["roofline"]; print(c, round(d["value"]), "MLUPS", round(d["ms_per_step"],3), "ms/step frac", round(r["frac"],3), {k:round(v,3) for k,v in r["frac_per_stage"].items()}, "parity", d["parity"] and (d["parity"].get("ok"), d["parity"].get("max_rel")), "e2e", d["e2e"] and (round(d["e2e"]["value"]), round(d["e2e"]["value_loop_only"])), "dropin", d["e2e_dropin"] and round(d["e2e_dropin"]["value"]), "cpu", d["cpu_baseline"] and (d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"]), d["clocks"])
    except Exception as e:
        print(c, "parse failed", e)
P
