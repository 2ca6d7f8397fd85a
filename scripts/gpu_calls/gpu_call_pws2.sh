#!/bin/bash
set -u
out=gpurun_out/r2pws
mkdir -p "$out"
export PYTHONUNBUFFERED=1
PRE='import sys; sys.path.insert(0, "scripts"); sys.argv=["x","none"]; import dev_time_r2 as D, torch; '
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pws_kernel -s 3 -c 1 -o /tmp/pws python -c "$PRE D.scalar_case((64,512,512), torch.float64, 1, True, tune={'variant': 30})" > "$out/ncu_pws.log" 2>&1
echo "ncu exit $?"; tail -3 "$out/ncu_pws.log"
ncu -i /tmp/pws.ncu-rep --page details --csv > "$out/pws_details.csv" 2>/dev/null
ncu -i /tmp/pws.ncu-rep --page raw --csv > "$out/pws_raw.csv" 2>/dev/null
ncu -i /tmp/pws.ncu-rep --page source --csv 2>/dev/null | gzip > "$out/pws_source.csv.gz"
ls -la "$out"
