#!/bin/bash
# fallback check: an allocator whose blocks cannot be IPC-exported (expandable segments) -> every rank falls back to the NCCL segments
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True timeout 400 $TR bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r3ag_bench_fallback.json 2> gpurun_out/r3ag_bench_fallback.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r3ag_bench_fallback.json")); print(round(d["value"],1), round(d["ms_per_step"],4), d["detail"]["launch"][:90])
except Exception as e:
    print("ERR", e)
print(open("gpurun_out/r3ag_bench_fallback.err").read()[-1200:])
PY
