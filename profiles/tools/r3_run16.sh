#!/bin/bash
# final verification: whole GPU suite on 2 GPUs (multi-GPU tests included), smoke, default bench line, reference arm
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r3r_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r3r_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3r_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r3r_smoke.log
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py > gpurun_out/r3r_bench.json 2> gpurun_out/r3r_bench.err; echo "bench rc=$?"
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --impl reference > gpurun_out/r3r_reference.json 2> gpurun_out/r3r_reference.err; echo "reference rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r3r_bench.json")); print("bench", round(d["value"],1), d["ms_per_step"], "e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],4), "launches", d["gpu_launches"], d.get("clocks"))
for k,v in d.get("also",{}).items():
    print(" also", k, {kk: (round(vv,2) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("value","ms_per_step","unit")}, "e2e", v.get("e2e",{}).get("value"))
print("cpu_baseline", d.get("cpu_baseline"))
r=json.load(open("gpurun_out/r3r_reference.json")); print("reference", r["value"], r.get("cpu_baseline"))
PY
