#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for cfg in "-1 1" "0 1" "-1 0"; do
set -- $cfg
PC_GRAPH_PRIORITY=$1 PC_PACK_STREAM=$2 timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu --no-also > gpurun_out/r2ah_p$1_s$2.json 2> gpurun_out/r2ah_p$1_s$2.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2ah_*.json")):
    d=json.load(open(f)); print(f, round(d["value"]), round(d["ms_per_step"],4))
PY
