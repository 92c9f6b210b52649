#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_r2.py -m gpu -q --timeout 300 -p no:cacheprovider -k "run_batches" > gpurun_out/r3ad_tests.log 2>&1; echo "tests rc=$?"
grep -v "^    " gpurun_out/r3ad_tests.log | tail -15
