#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
export HB_LAYERS=0
CMD="python profiles/tools/halo_bench.py"
timeout 300 $CMD > gpurun_out/r2d_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 6 -c 4 -o gpurun_out/r2d_halo_full $CMD > gpurun_out/r2d_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2d_ncu.log
