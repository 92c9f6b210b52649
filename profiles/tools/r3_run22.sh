#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for V in 4 3 4 3; do
PC_BN_REDUCE_BPS=$V timeout 600 python bench.py --steps 100 --warmup 5 --no-also --no-cpu > gpurun_out/r3z_deep_$V.json 2> gpurun_out/r3z_deep_$V.err; echo "deep bps=$V rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r3z_deep_$V.json")); print("bps=$V", round(d["value"],1), round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["roofline"]["by_entry_point_ms"].items() if "bwd_reduce" in k})
PY
done
