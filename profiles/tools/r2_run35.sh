#!/bin/bash
# compute-sanitizer memcheck over the kernels written / rewritten in round 2 (small cases)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1700 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_halo.py tests/test_gpu_r2.py tests/test_gpu_parity.py -m gpu -q --timeout 1500 -p no:cacheprovider -x \
  -k "(halo_wgrad and (case0 or case2 or case4 or case5)) or (halo_1x1 and (case0 or case4)) or (stride2 and case3) or (stem and (shape1 or shape0)) or (mfcc_vs_reference and 4000) or fused_views or (test_halo_forward and case6 and 1) or (halo_dgrad and case6)" > gpurun_out/r2ag_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|passed|failed|error" gpurun_out/r2ag_memcheck.log | head -20
tail -5 gpurun_out/r2ag_memcheck.log | cut -c1-200
