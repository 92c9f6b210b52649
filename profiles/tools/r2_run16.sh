#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_r2.py -m gpu -q --timeout 300 -p no:cacheprovider -k "stem" > gpurun_out/r2p_tests.log 2>&1; echo "stem tests rc=$?"
tail -3 gpurun_out/r2p_tests.log
timeout 300 python profiles/tools/stem_bench.py 2>&1 | tee gpurun_out/r2p_stem_bench.log
