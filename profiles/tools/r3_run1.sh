#!/bin/bash
# Peer-memory exchange: single-device kernel tests, 2-GPU tests (IPC mapping, one-graph step vs NCCL segments vs oracle replicas),
# then the 2-GPU bench lines in both exchange modes.
set -u
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_peer.py -m gpu -q --timeout 120 -p no:cacheprovider > gpurun_out/r3a_peer.log 2>&1; echo "peer tests rc=$?"
tail -15 gpurun_out/r3a_peer.log
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/r3a_multi.log 2>&1; echo "multi tests rc=$?"
tail -30 gpurun_out/r3a_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for MODE in peer nccl; do
  PC_DP_EXCHANGE=$MODE timeout 600 $TR bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r3a_bench_${MODE}_n$N.json 2> gpurun_out/r3a_bench_${MODE}_n$N.err; echo "bench $MODE rc=$?"
  PC_DP_EXCHANGE=$MODE timeout 600 $TR bench.py --gpus $N --workload supcon_8192 --steps 50 --warmup 5 > gpurun_out/r3a_supcon_${MODE}_n$N.json 2> gpurun_out/r3a_supcon_${MODE}_n$N.err; echo "supcon $MODE rc=$?"
done
python - <<PY
import json
for m in ["peer","nccl"]:
  for f in [f"r3a_bench_{m}_n$N",f"r3a_supcon_{m}_n$N"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],1), round(d["ms_per_step"],4), d.get("detail",{}).get("launch","")[:60], d["config"].get("exchange",""))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
