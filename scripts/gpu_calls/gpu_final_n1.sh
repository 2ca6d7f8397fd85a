#!/bin/bash
# Final single-GPU measurements of the round: tests, bench lines for every configuration,
# launch list and ncu captures (exported as CSV; the .ncu-rep files are too large to keep).
set -u
out=gpurun_out/r2final
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -m gpu -x > "$out/pytest_gpu.log" 2>&1
echo "pytest exit $?"; tail -5 "$out/pytest_gpu.log"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 > "$out/bench_c3_n1.json" 2> "$out/bench_c3_n1.err"; echo "bench c3 exit $?"
for c in c2 c4 c5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 3 > "$out/bench_${c}_n1.json" 2> "$out/bench_${c}_n1.err"; echo "bench $c exit $?"
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > "$out/bench_reference_c3.json" 2> "$out/bench_reference_c3.err"; echo "reference exit $?"
python - <<P
import json
for c in ("c3","c2","c4","c5"):
    try:
        d=json.loads([l for l in open("$out/bench_%s_n1.json"%c).read().strip().splitlines() if l.startswith("{")][-1])
        r=d["roofline"]; print(c, round(d["value"]), "MLUPS", round(d["ms_per_step"],3), "ms/step frac", round(r["frac"],3), {k:round(v,3) for k,v in r["frac_per_stage"].items()}, "parity", d["parity"] and (d["parity"].get("ok"), d["parity"].get("max_rel")), "e2e", d["e2e"] and (round(d["e2e"]["value"]), round(d["e2e"]["value_loop_only"])), "dropin", d["e2e_dropin"] and round(d["e2e_dropin"]["value"]), "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], "launches", d["gpu_launches"])
        if "whole_stage" in r: print("   whole stage", r["whole_stage"], r["scalar_stage_ms_per_launch"])
    except Exception as e:
        print(c, "parse failed", e)
try:
    d=json.loads(open("$out/bench_reference_c3.json").read().strip().splitlines()[-1]); print("reference", d["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["seconds_min_max"])
except Exception as e: print("reference parse failed", e)
P
# launch list (cold-cache, serialised: compare shares)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$out/launches_bench_c3.csv" \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$out/launches_bench_c4.csv" \
    python bench.py --config c4 --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity > /dev/null 2>&1
prof() {
  local name=$1 regex=$2 skip=$3 cmd=$4
  timeout 300 ncu --set full --clock-control none -k regex:$regex -s $skip -c 1 -o "/tmp/$name" $cmd > "$out/ncu_$name.log" 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > "$out/ncu_${name}_raw.csv" 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page details --csv > "$out/ncu_${name}_details.csv" 2>/dev/null
}
prof c3_stage1 fused_stage_ws 16 "python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity --no-graph"
prof c3_stage0 fused_stage_ws 15 "python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity --no-graph"
prof c4_tensor tensor_stage_ws 16 "python bench.py --config c4 --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity --no-graph"
prof c4_scalar fused_stage_ws 16 "python bench.py --config c4 --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity --no-graph"
prof c5_stage1 fused_stage_ws 16 "python bench.py --config c5 --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity --no-graph"
grep -h "dram__bytes_read.sum\|dram__bytes_write.sum" /dev/null
python - <<P
import csv
for n in ("c3_stage1","c3_stage0","c4_tensor","c4_scalar","c5_stage1"):
    try:
        rows=list(csv.reader(open("$out/ncu_%s_raw.csv"%n))); h=rows[0]; u=rows[1]; v=rows[2]
        g=lambda k:[(x,y) for a,x,y in zip(h,u,v) if a==k][0]
        print(n, v[h.index("Kernel Name")][:60], "dur", g("gpu__time_duration.sum"), "rd", g("dram__bytes_read.sum"), "wr", g("dram__bytes_write.sum"))
    except Exception as e: print(n, "failed", e)
P
du -sh "$out"
