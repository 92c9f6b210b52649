#!/bin/bash
# N GPUs: multi-GPU tests (N>=2) + scaling lines for the main workload, configs[4] and configs[3]
set -u
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2fin_multi.log 2>&1; echo "multi tests rc=$?"
  tail -4 gpurun_out/r2fin_multi.log
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r2fin_bench_n$N.json 2> gpurun_out/r2fin_bench_n$N.err; echo "bench rc=$?"
timeout 600 $TR bench.py --gpus $N --workload supcon_8192 --steps 50 --warmup 5 > gpurun_out/r2fin_supcon_n$N.json 2> gpurun_out/r2fin_supcon_n$N.err; echo "supcon rc=$?"
timeout 600 $TR bench.py --gpus $N --workload train_cnn_deep_4096 --steps 30 --warmup 5 > gpurun_out/r2fin_bench4096_n$N.json 2> gpurun_out/r2fin_bench4096_n$N.err; echo "bench 4096 rc=$?"
python - <<PY
import json
for f in ["r2fin_bench_n$N","r2fin_supcon_n$N","r2fin_bench4096_n$N"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],1), round(d["ms_per_step"],4))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
