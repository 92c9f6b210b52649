#!/bin/bash
# N-GPU scaling lines with the peer-memory exchanges (main workload + SupCon 8192)
set -u
N=${1:-8}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r3t_bench_n$N.json 2> gpurun_out/r3t_bench_n$N.err; echo "bench rc=$?"
timeout 300 $TR bench.py --gpus $N --workload supcon_8192 --steps 50 --warmup 5 > gpurun_out/r3t_supcon_n$N.json 2> gpurun_out/r3t_supcon_n$N.err; echo "supcon rc=$?"
python - <<PY
import json
for f in ["r3t_bench_n$N","r3t_supcon_n$N"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],1), round(d["ms_per_step"],4), d.get("detail",{}).get("launch","")[:60], d["config"].get("exchange",""))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
grep -i "warn\|fail\|error" gpurun_out/r3t_bench_n$N.err | head -5
