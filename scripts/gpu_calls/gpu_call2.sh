#!/bin/bash
set -u
out=gpurun_out/r2c2
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -q -m gpu -x > "$out/pytest_gpu.log" 2>&1
echo "pytest exit $?"; tail -15 "$out/pytest_gpu.log"
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/mb scripts/microbench_rw.cu && timeout 120 /tmp/mb > "$out/microbench_rw.log" 2>&1
cat "$out/microbench_rw.log"
timeout 900 python scripts/dev_time_r2.py > "$out/time_r2.log" 2>&1
cat "$out/time_r2.log"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > "$out/bench_n1.json" 2> "$out/bench_n1.err"
tail -c 1500 "$out/bench_n1.json"; tail -5 "$out/bench_n1.err"
