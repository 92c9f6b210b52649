#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_r2.py -m gpu -q --timeout 300 -p no:cacheprovider -k "stem_forward" > gpurun_out/r2j_tests.log 2>&1; echo "stem fwd tests rc=$?"
tail -30 gpurun_out/r2j_tests.log
