#!/bin/bash
# Last single-GPU pass of the round on the final tree: full GPU suite, smoke, c3 / c4 bench
# lines, launch lists, ncu captures of the two kernels that changed last (stage 0 with the
# quadratic path, config 4's generated scalar stage).
set -u
out=gpurun_out/r2final3
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -m gpu > "$out/pytest_gpu.log" 2>&1
echo "pytest exit $?"; tail -6 "$out/pytest_gpu.log"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for c in c3 c4; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 3 > "$out/bench_${c}_n1.json" 2> "$out/bench_${c}_n1.err"; echo "bench $c exit $?"
done
python - <<P
import json
for c in ("c3","c4"):
    try:
        d=json.loads([l for l in open("$out/bench_%s_n1.json"%c).read().strip().splitlines() if l.startswith("{")][-1])
        r=d["roofline"]; print(c, round(d["value"]), "MLUPS", round(d["ms_per_step"],3), "ms/step frac", round(r["frac"],3), {k:round(v,3) for k,v in r["frac_per_stage"].items()}, "parity", d["parity"] and (d["parity"].get("ok"), d["parity"].get("max_rel")), "e2e", d["e2e"] and (round(d["e2e"]["value"]), round(d["e2e"]["value_loop_only"])), "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], "launches", d["gpu_launches"])
        dr=d.get("e2e_dropin")
        if dr: print("   dropin", round(dr["value"]), "loop only", round(dr["value_loop_only"]), "h2d ms", round(dr["h2d_ms"],1), "loop ms/step", round(dr["loop_ms_per_step"],3), "d2h ms", round(dr["d2h_ms"],1))
        if "whole_stage" in r: print("   whole stage", r["whole_stage"], r["scalar_stage_ms_per_launch"])
    except Exception as e:
        print(c, "parse failed", e)
P
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
prof c3_stage0_quadratic fused_stage_ws 15 "python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity --no-graph"
prof c4_scalar_generated fused_stage_ws 16 "python bench.py --config c4 --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity --no-graph"
python - <<P
import csv
for n in ("c3_stage0_quadratic","c4_scalar_generated"):
    try:
        rows=list(csv.reader(open("$out/ncu_%s_raw.csv"%n))); h=rows[0]; u=rows[1]; v=rows[2]
        g=lambda k:[(x,y) for a,x,y in zip(h,u,v) if a==k][0]
        print(n, v[h.index("Kernel Name")][:60], "dur", g("gpu__time_duration.sum"), "rd", g("dram__bytes_read.sum"), "wr", g("dram__bytes_write.sum"))
    except Exception as e: print(n, "failed", e)
P
head -c 1500 "$out/launches_bench_c3.csv" | tail -c 700
du -sh "$out"
