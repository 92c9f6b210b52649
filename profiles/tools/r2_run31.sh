#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu --no-also > gpurun_out/r2ad_bench.json 2> gpurun_out/r2ad_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2ad_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
print(d["roofline"]["by_entry_point_ms"])
PY
