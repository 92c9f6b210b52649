#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_r2.py -m gpu -q --timeout 300 -p no:cacheprovider -k "stem" > gpurun_out/r2n_tests.log 2>&1; echo "stem tests rc=$?"
tail -3 gpurun_out/r2n_tests.log
CMD="python profiles/tools/stem_bench.py"
timeout 300 $CMD > gpurun_out/r2n_stem_bench.log 2>&1 && cat gpurun_out/r2n_stem_bench.log && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stem_fwd_kernel|stem_bwd_pool_kernel|stem_gram_kernel" -s 6 -c 3 -o gpurun_out/r2n_stem_full $CMD > gpurun_out/r2n_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2n_ncu.log
