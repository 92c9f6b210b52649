#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD="python bench.py --workload frontend --steps 2 --warmup 1 --no-cpu"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"frontend_kernel" -s 2 -c 1 -o gpurun_out/r2w_fe_full $CMD > gpurun_out/r2w_ncu.log 2>&1; echo "ncu rc=$?"
