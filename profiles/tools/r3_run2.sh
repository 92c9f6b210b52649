#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_peer.py -m gpu -q --timeout 120 -p no:cacheprovider > gpurun_out/r3c_peer.log 2>&1; echo "peer tests rc=$?"
tail -5 gpurun_out/r3c_peer.log
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/r3c_multi.log 2>&1; echo "multi tests rc=$?"
tail -5 gpurun_out/r3c_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r3c_bench_n$N.json 2> gpurun_out/r3c_bench_n$N.err; echo "bench rc=$?"
timeout 600 $TR bench.py --gpus $N --workload train_cnn_deep_4096 --steps 30 --warmup 5 > gpurun_out/r3c_bench4096_n$N.json 2> gpurun_out/r3c_bench4096_n$N.err; echo "bench4096 rc=$?"
timeout 600 python bench.py --steps 50 --warmup 5 --no-also --no-cpu > gpurun_out/r3c_bench_n1.json 2> gpurun_out/r3c_bench_n1.err; echo "bench n1 rc=$?"
python - <<PY
import json
for f in ["r3c_bench_n$N","r3c_bench4096_n$N","r3c_bench_n1"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],1), round(d["ms_per_step"],4), d.get("detail",{}).get("launch","")[:60])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
