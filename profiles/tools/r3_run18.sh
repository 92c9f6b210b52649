#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD="python bench.py --workload supcon_8192 --steps 1 --warmup 3 --no-cpu --no-also"
timeout 300 $CMD > gpurun_out/r3u_supcon_plain.json 2> gpurun_out/r3u_supcon_plain.err; echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"supcon|pack_rows" -s 12 -c 8 -o /tmp/r3u_supcon_full $CMD > gpurun_out/r3u_supcon_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r3u_supcon_full.ncu-rep --page raw --csv > gpurun_out/r3u_supcon_full_raw.csv 2>/dev/null
ls -la gpurun_out/r3u_*; tail -3 gpurun_out/r3u_supcon_ncu.log
