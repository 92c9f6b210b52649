#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2f_multi.log 2>&1; echo "multi tests rc=$?"
tail -15 gpurun_out/r2f_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err; echo "bench n2 rc=$?"
timeout 600 python bench.py --workload supcon_8192 --steps 50 --warmup 5 --no-cpu > gpurun_out/r2f_supcon_n1.json 2> gpurun_out/r2f_supcon_n1.err; echo "supcon n1 rc=$?"
timeout 600 $TR bench.py --gpus 2 --workload supcon_8192 --steps 50 --warmup 5 > gpurun_out/r2f_supcon_n2.json 2> gpurun_out/r2f_supcon_n2.err; echo "supcon n2 rc=$?"
timeout 600 $TR bench.py --gpus 2 --workload train_cnn_deep_4096 --steps 30 --warmup 5 > gpurun_out/r2f_bench4096_n2.json 2> gpurun_out/r2f_bench4096_n2.err; echo "bench 4096 n2 rc=$?"
python - <<'PY'
import json
for f in ["r2f_bench_n2","r2f_supcon_n1","r2f_supcon_n2","r2f_bench4096_n2"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["ms_per_step"], d.get("config"))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
