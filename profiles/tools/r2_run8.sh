#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_r2.py -m gpu -q --timeout 600 -p no:cacheprovider -k "stem" > gpurun_out/r2h_stem_tests.log 2>&1; echo "stem tests rc=$?"
tail -25 gpurun_out/r2h_stem_tests.log
