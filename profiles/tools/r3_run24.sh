#!/bin/bash
# final state: whole GPU suite, smoke, default bench line, reference arm
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 -p no:cacheprovider > gpurun_out/r3ac_tests.log 2>&1; echo "tests rc=$?"
tail -2 gpurun_out/r3ac_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3ac_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r3ac_smoke.log
timeout 900 python bench.py > gpurun_out/r3ac_bench.json 2> gpurun_out/r3ac_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r3ac_reference.json 2> gpurun_out/r3ac_reference.err; echo "reference rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r3ac_bench.json")); print("bench", round(d["value"],1), d["ms_per_step"], "e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],4), "launches", d["gpu_launches"], d.get("clocks"))
for k,v in d.get("also",{}).items():
    if isinstance(v,dict):
        e=v.get("e2e"); print(" also", k, round(v.get("value",0),1), v.get("ms_per_step"), e if not isinstance(e,dict) else e.get("value"))
r=json.load(open("gpurun_out/r3ac_reference.json")); print("reference", r["value"], r.get("steps"))
PY
