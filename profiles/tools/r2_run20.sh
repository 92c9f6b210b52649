#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD="python profiles/tools/stem_bench.py"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stem_bwd_pool_kernel" -s 3 -c 1 -o gpurun_out/r2t_stem_bwd_full $CMD > gpurun_out/r2t_ncu2.log 2>&1; echo "ncu rc=$?"
