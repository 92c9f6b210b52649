#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in 1 0; do
PC_WGRAD_STREAM=$v timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu --no-also > gpurun_out/r2af_bench_ws$v.json 2> gpurun_out/r2af_bench_ws$v.err
done
python - <<'PY'
import json
for v in (1,0):
    d=json.load(open(f"gpurun_out/r2af_bench_ws{v}.json"))
    print("PC_WGRAD_STREAM",v, d["value"], d["ms_per_step"])
PY
