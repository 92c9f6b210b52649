#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_halo.py -m gpu -q --timeout 120 -p no:cacheprovider -k "halo_wgrad" > gpurun_out/r2y_tests.log 2>&1; echo "wgrad tests rc=$?"
tail -30 gpurun_out/r2y_tests.log | cut -c1-220
