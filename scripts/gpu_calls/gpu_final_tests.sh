#!/bin/bash
set -u
out=gpurun_out/r2tests
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -q -m gpu -x > "$out/pytest_gpu.log" 2>&1
echo "pytest exit $?"; tail -6 "$out/pytest_gpu.log"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
