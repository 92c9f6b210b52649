#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_r2.py tests/test_reference_suite.py -m gpu -q --timeout 300 -p no:cacheprovider -k "dataset or fused or trainer or reference or transform" > gpurun_out/r2ac_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2ac_tests.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/r2ac_bench.json 2> gpurun_out/r2ac_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2ac_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
for k,v in d.get("also",{}).items():
    print(k, v["value"] if isinstance(v,dict) else v, v.get("ms_per_step") if isinstance(v,dict) else "")
PY
