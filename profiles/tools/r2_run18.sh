#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_halo.py -m gpu -q --timeout 300 -p no:cacheprovider -k "1x1" > gpurun_out/r2r_tests.log 2>&1; echo "1x1 tests rc=$?"
tail -5 gpurun_out/r2r_tests.log
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-also"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2r_launches.csv $CMD > gpurun_out/r2r_ncu.log 2>&1; echo "ncu rc=$?"
