#!/bin/bash
set -u
N=2
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for BL in 0 16 148; do
PC_PEER_AR_BLOCKS=$BL timeout 400 $TR bench.py --gpus $N --steps 60 --warmup 5 > gpurun_out/r3q_bench_b$BL.json 2> gpurun_out/r3q_bench_b$BL.err; echo "bench $BL rc=$?"
done
timeout 400 $TR bench.py --gpus $N --workload train_cnn_small --steps 100 --warmup 5 > gpurun_out/r3q_small_n2.json 2> gpurun_out/r3q_small_n2.err; echo "small rc=$?"
python - <<PY
import json
for f in ["r3q_bench_b0","r3q_bench_b16","r3q_bench_b148","r3q_small_n2"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
