#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_peer.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r3ab_multi.log 2>&1; echo "multi tests rc=$?"
tail -3 gpurun_out/r3ab_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 2 --steps 60 --warmup 5 > gpurun_out/r3ab_bench_n2.json 2> gpurun_out/r3ab_bench_n2.err; echo "bench rc=$?"
timeout 400 $TR bench.py --gpus 2 --workload train_cnn_small --steps 100 --warmup 5 > gpurun_out/r3ab_small_n2.json 2> gpurun_out/r3ab_small_n2.err; echo "small rc=$?"
python - <<PY
import json
for f in ["r3ab_bench_n2","r3ab_small_n2"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
