#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for CFG in "1 0" "1 -1" "3 -1" "3 0" "6 -1"; do
set -- $CFG
PC_WGRAD_HALO_SPLIT_MUL=$1 PC_GRAPH_PRIORITY=$2 timeout 600 python bench.py --steps 100 --warmup 5 --no-also --no-cpu > gpurun_out/r3w_m$1_p$2.json 2> gpurun_out/r3w_m$1_p$2.err; echo "mul=$1 prio=$2 rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r3w_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["value"],1), round(d["ms_per_step"],4))
    except Exception as e:
        print(f, "ERR", e)
PY
