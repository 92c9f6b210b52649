#!/bin/bash
# Round-2 GPU call 1: full GPU test suite, baseline bench lines (own arm, torch_cuda yardstick), bf16 diagnostic,
# ncu launch lists for the front end and a cnn_small step, one full capture of frontend_kernel.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl torch_cuda --steps 20 --warmup 5 > gpurun_out/r2a_torch_cuda_deep.json 2> gpurun_out/r2a_torch_cuda_deep.err; echo "tc deep rc=$?"
timeout 300 python bench.py --impl torch_cuda --workload train_cnn_small --steps 50 --warmup 5 > gpurun_out/r2a_torch_cuda_small.json 2> gpurun_out/r2a_torch_cuda_small.err; echo "tc small rc=$?"
timeout 300 python bench.py --impl torch_cuda --workload supcon_8192 --steps 10 --warmup 3 > gpurun_out/r2a_torch_cuda_supcon.json 2> gpurun_out/r2a_torch_cuda_supcon.err; echo "tc supcon rc=$?"
timeout 300 python bench.py --impl torch_cuda --workload frontend --steps 5 --warmup 3 > gpurun_out/r2a_torch_cuda_fe.json 2> gpurun_out/r2a_torch_cuda_fe.err; echo "tc fe rc=$?"
timeout 300 python profiles/tools/bf16_diag.py > gpurun_out/r2a_bf16.log 2>&1; echo "bf16 rc=$?"
# ncu: launch lists (front end; cnn_small eager step), then one full capture of the front-end kernel
FE="python bench.py --workload frontend --steps 1 --warmup 3 --no-also --no-cpu"
SM="python bench.py --workload train_cnn_small --steps 2 --warmup 3 --no-also --no-cpu --no-graph"
timeout 300 $FE > gpurun_out/r2a_fe_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2a_fe_launches.csv $FE > gpurun_out/r2a_fe_ncu.log 2>&1; echo "ncu fe rc=$?"
timeout 300 $SM > gpurun_out/r2a_small_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 400 --csv --log-file gpurun_out/r2a_small_launches.csv $SM > gpurun_out/r2a_small_ncu.log 2>&1; echo "ncu small rc=$?"
timeout 300 $FE > gpurun_out/r2a_fe_plain2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:frontend_kernel -s 3 -c 1 -o gpurun_out/r2a_fe_full $FE > gpurun_out/r2a_fe_full.log 2>&1; echo "ncu fe full rc=$?"
ls -la gpurun_out | tail -20
