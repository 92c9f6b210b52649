#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD="python profiles/tools/stem_bench.py"
timeout 300 $CMD > gpurun_out/r2k_stem_bench.log 2>&1 && cat gpurun_out/r2k_stem_bench.log && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stem_fwd_kernel|stem_bwd_pool_kernel" -s 4 -c 2 -o gpurun_out/r2k_stem_full $CMD > gpurun_out/r2k_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2k_ncu.log
