#!/bin/bash
# final bench lines of round 2 (N = 1): product arm, reference arm, stock-PyTorch-on-GPU yardstick, secondary workloads
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench rc=$?"
timeout 1500 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_reference_arm.json 2> gpurun_out/r2_final_reference_arm.err; echo "ref rc=$?"
timeout 900 python bench.py --impl torch_cuda --steps 20 --warmup 3 > gpurun_out/r2_final_torch_cuda_deep.json 2> gpurun_out/r2_final_torch_cuda_deep.err; echo "tc rc=$?"
timeout 900 python bench.py --workload supcon_8192 --steps 50 --warmup 5 --no-cpu > gpurun_out/r2_final_supcon_8192.json 2> gpurun_out/r2_final_supcon.err; echo "supcon rc=$?"
timeout 900 python bench.py --workload train_cnn_small --steps 100 --warmup 10 > gpurun_out/r2_final_cnn_small.json 2> gpurun_out/r2_final_small.err; echo "small rc=$?"
timeout 900 python bench.py --workload frontend --steps 10 --warmup 3 > gpurun_out/r2_final_frontend.json 2> gpurun_out/r2_final_fe.err; echo "fe rc=$?"
python - <<'PY'
import json
for f in ["r2_final_bench","r2_final_reference_arm","r2_final_torch_cuda_deep","r2_final_supcon_8192","r2_final_cnn_small","r2_final_frontend"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, d.get("value"), d.get("unit"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("kind"))
    except Exception as e: print(f, "ERR", e)
PY
