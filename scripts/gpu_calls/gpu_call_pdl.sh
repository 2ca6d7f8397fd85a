#!/bin/bash
# A/B of programmatic dependent launch (tune pdl) on one GPU: bit-identity test, then bench lines
mkdir -p gpurun_out/pdl
timeout 400 python -m pytest tests/test_gpu_config_parity.py -k programmatic -x -q 2>&1 | tail -5
echo "pytest pdl exit $?"
run() {
  tag=$1; shift
  for pdl in 0 1 0 1; do
    timeout 300 python bench.py "$@" --pdl $pdl --no-cpu --no-e2e --steps 20 --warmup 3 \
      > gpurun_out/pdl/${tag}_pdl${pdl}.json.tmp 2> gpurun_out/pdl/${tag}_pdl${pdl}.err
    python - gpurun_out/pdl/${tag}_pdl${pdl}.json.tmp "$tag" $pdl <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(sys.argv[2], "pdl", sys.argv[3], round(d["value"]), "MLUPS", round(d["ms_per_step"], 4),
          "ms/step frac", round(d["roofline"]["frac"], 4), "parity", d["parity"]["ok"],
          d["parity"]["max_rel"], "launches", d.get("gpu_launches"))
except Exception as e:
    print(sys.argv[2], "pdl", sys.argv[3], "FAILED", e)
PY
    cat gpurun_out/pdl/${tag}_pdl${pdl}.json.tmp >> gpurun_out/pdl/${tag}_pdl${pdl}.json
    rm -f gpurun_out/pdl/${tag}_pdl${pdl}.json.tmp
  done
}
run c3 --config c3
run c2 --config c2
run c3nx64 --config c3 --nx 64
run c5 --config c5
run c4 --config c4
tail -3 gpurun_out/pdl/*.err | tail -30
