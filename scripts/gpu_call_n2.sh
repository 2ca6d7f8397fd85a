#!/bin/bash
# gpurun --gpus 2 -- 'bash scripts/gpu_call_n2.sh'
set -u
out=gpurun_out/r2n2
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_distributed.py -q -m gpu -x > "$out/pytest_dist.log" 2>&1
echo "pytest dist exit $?"; tail -4 "$out/pytest_dist.log"
for c in c3; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --config $c --steps 20 --warmup 3 --no-e2e > "$out/bench_${c}_n2.json" 2> "$out/bench_${c}_n2.err"
  echo "bench $c N=2 exit $?"
  python - <<P
import json
try:
    d=json.loads([l for l in open("$out/bench_${c}_n2.json").read().strip().splitlines() if l.startswith("{")][-1])
    r=d["roofline"]; print("$c N=2", round(d["value"]), "MLUPS", round(d["ms_per_step"],3), "ms/step frac", round(r["frac"],3), {k:round(v,4) for k,v in r["ms_per_launch"].items()}, "parity", d["parity"] and (d["parity"]["ok"], d["parity"]["max_rel"]), "launches", d["gpu_launches"])
except Exception as e:
    print("$c parse failed", e)
P
  tail -2 "$out/bench_${c}_n2.err"
done
