#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_halo.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "32_channel" > gpurun_out/r3k_halo.log 2>&1; echo "c32 tests rc=$?"
tail -30 gpurun_out/r3k_halo.log
timeout 900 python -m pytest tests/test_gpu_halo.py tests/test_gpu_tc.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r3k_all.log 2>&1; echo "halo+tc tests rc=$?"
tail -5 gpurun_out/r3k_all.log
