#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_halo.py -m gpu -q --timeout 120 -p no:cacheprovider > gpurun_out/r2c_halo_tests.log 2>&1; echo "halo tests rc=$?"
tail -25 gpurun_out/r2c_halo_tests.log
timeout 300 python profiles/tools/halo_bench.py > gpurun_out/r2c_halo_bench.log 2>&1; echo "halo bench rc=$?"
cat gpurun_out/r2c_halo_bench.log
