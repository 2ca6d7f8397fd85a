#!/bin/bash
# First gpurun call of the next round: everything written after round 1's GPU budget ran
# out, plus the numbers needed to decide what to tune.  Run from the repo root:
#   gpurun --timeout 1500 -- 'bash scripts/round2_first_call.sh'
# Outputs land in gpurun_out/round2/.
set -u
out=gpurun_out/round2
mkdir -p "$out"
export PYTHONUNBUFFERED=1

echo "== validated GPU suite ==" | tee "$out/summary.txt"
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_zz_gpu_round2_candidates.py \
    > "$out/pytest_gpu.log" 2>&1
echo "exit $?" | tee -a "$out/summary.txt"; tail -3 "$out/pytest_gpu.log" | tee -a "$out/summary.txt"

echo "== round-2 candidates (not yet run on hardware) ==" | tee -a "$out/summary.txt"
timeout 900 python -m pytest tests/test_zz_gpu_round2_candidates.py -q -m gpu \
    > "$out/pytest_candidates.log" 2>&1
echo "exit $?" | tee -a "$out/summary.txt"; tail -15 "$out/pytest_candidates.log" | tee -a "$out/summary.txt"

echo "== bench N=1 ==" | tee -a "$out/summary.txt"
timeout 600 python bench.py --steps 20 --warmup 3 > "$out/bench_n1.json" 2> "$out/bench_n1.err"
tail -c 600 "$out/bench_n1.json" | tee -a "$out/summary.txt"

echo "== stand-alone kernels, fused GW vs unfused, generated stage vs hand-written ==" | tee -a "$out/summary.txt"
timeout 600 python scripts/dev_time_api.py > "$out/time_api.log" 2>&1
echo "-- with PSB_FD_PIPELINE=1 PSB_REDUCE_V2=1" >> "$out/time_api.log"
PSB_FD_PIPELINE=1 PSB_REDUCE_V2=1 timeout 600 python scripts/dev_time_api.py >> "$out/time_api.log" 2>&1
timeout 600 python scripts/dev_time_round2.py > "$out/time_round2.log" 2>&1
cat "$out/time_api.log" "$out/time_round2.log" | tee -a "$out/summary.txt"

echo "== example: golden constraint ==" | tee -a "$out/summary.txt"
timeout 300 python examples/scalar_preheating.py -grid 32 32 32 --end-time 1 > "$out/example32.log" 2>&1
tail -4 "$out/example32.log" | tee -a "$out/summary.txt"
