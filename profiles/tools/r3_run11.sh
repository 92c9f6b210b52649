#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_halo.py -m gpu -q --timeout 300 -p no:cacheprovider -k "wgrad" > gpurun_out/r3l_halo.log 2>&1; echo "wgrad tests rc=$?"
grep -v "^E    \|^    " gpurun_out/r3l_halo.log | tail -30
