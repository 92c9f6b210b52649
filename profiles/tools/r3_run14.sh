#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for V in 4 2 1; do
PC_BN_REDUCE_BPS=$V timeout 600 python bench.py --workload train_cnn_small --steps 200 --warmup 10 --no-also --no-cpu > gpurun_out/r3p_small_$V.json 2> gpurun_out/r3p_small_$V.err; echo "small $V rc=$?"
PC_BN_REDUCE_BPS=$V timeout 600 python bench.py --steps 100 --warmup 5 --no-also --no-cpu > gpurun_out/r3p_deep_$V.json 2> gpurun_out/r3p_deep_$V.err; echo "deep $V rc=$?"
done
python - <<PY
import json
for V in [4,2,1]:
  for f in [f"r3p_small_{V}",f"r3p_deep_{V}"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],1), round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["roofline"]["by_entry_point_ms"].items() if "bwd_reduce" in k})
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
