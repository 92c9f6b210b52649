#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/pc_probe_bin profiles/tools/probe_umma_tma.cu > gpurun_out/r2b_probe_build.log 2>&1
timeout 120 /tmp/pc_probe_bin > gpurun_out/r2b_probe.log 2>&1; echo "probe rc=$?"
cat gpurun_out/r2b_probe.log
timeout 600 python -m pytest tests/test_gpu_r2.py tests/test_reference_suite.py tests/test_gpu_tc.py -m gpu -q --timeout 600 -p no:cacheprovider -k "graph_replay or checkpoint or tiny_dy or reference_test or bf16" > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2b_tests.log
timeout 600 python profiles/tools/stem_grad_diag.py > gpurun_out/r2b_stem_diag.log 2>&1; echo "diag rc=$?"
grep -v Warning gpurun_out/r2b_stem_diag.log | tail -5
