#!/bin/bash
# gpurun --gpus 2 -- 'bash scripts/gpu_calls/gpu_call_n2b.sh': bench lines of the final tree on 2 GPUs
# (in-bench parity across the rank boundary), plus the same with programmatic dependent launch
set -u
out=gpurun_out/r2n2b
mkdir -p "$out"
export PYTHONUNBUFFERED=1
run() {
  tag=$1; shift
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-e2e "$@" \
    > "$out/$tag.json" 2> "$out/$tag.err"
  echo "$tag exit $?"
  python - "$out/$tag.json" "$tag" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(sys.argv[2], round(d["value"]), "MLUPS", round(d["ms_per_step"], 4), "ms/step parity",
          d["parity"]["ok"], d["parity"]["max_rel"], "launches", d.get("gpu_launches"))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run bench_c3_n2 --config c3
run bench_c3_n2_pdl --config c3 --pdl 1
run bench_c4_n2 --config c4
tail -n 3 "$out"/*.err | tail -n 20
