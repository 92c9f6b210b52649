#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/r2u_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r2u_tests.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2u_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
print(d["roofline"]["by_entry_point_ms"])
for k,v in d.get("also",{}).items():
    print(k, v["value"] if isinstance(v,dict) else v, v.get("ms_per_step") if isinstance(v,dict) else "")
PY
