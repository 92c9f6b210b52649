#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "sync_batchnorm" > gpurun_out/r3e_multi.log 2>&1; echo "syncbn tests rc=$?"
tail -40 gpurun_out/r3e_multi.log
