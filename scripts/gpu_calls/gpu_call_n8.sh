#!/bin/bash
# gpurun --gpus 8 -- 'bash scripts/gpu_calls/gpu_call_n8.sh'
set -u
out=gpurun_out/r2n8
mkdir -p "$out"
export PYTHONUNBUFFERED=1
nvidia-smi -L | wc -l
for c in c3 c4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
     bench.py --gpus 8 --config $c --steps 20 --warmup 3 --no-e2e > "$out/bench_${c}_n8.json" 2> "$out/bench_${c}_n8.err"
  echo "bench $c N=8 exit $?"
  python - <<P
import json
try:
    d=json.loads([l for l in open("$out/bench_${c}_n8.json").read().strip().splitlines() if l.startswith("{")][-1])
    r=d["roofline"]; print("$c N=8", round(d["value"]), "MLUPS", round(d["ms_per_step"],4), "ms/step frac", round(r["frac"],3), {k:round(v,4) for k,v in r["ms_per_launch"].items()}, "parity", d["parity"] and (d["parity"]["ok"], d["parity"]["max_rel"]))
except Exception as e:
    print("$c parse failed", e)
P
  tail -3 "$out/bench_${c}_n8.err"
done
