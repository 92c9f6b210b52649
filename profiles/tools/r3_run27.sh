#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_peer.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r3ah_multi.log 2>&1; echo "multi tests rc=$?"
tail -2 gpurun_out/r3ah_multi.log
