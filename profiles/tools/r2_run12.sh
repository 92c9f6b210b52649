#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_r2.py -m gpu -q --timeout 300 -p no:cacheprovider -k "stem" > gpurun_out/r2l_tests.log 2>&1; echo "stem tests rc=$?"
tail -5 gpurun_out/r2l_tests.log
timeout 300 python profiles/tools/stem_bench.py 2>&1 | tee gpurun_out/r2l_stem_bench.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2l_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
print(d["roofline"]["by_entry_point_ms"])
for k,v in d.get("also",{}).items():
    print(k, v["value"] if isinstance(v,dict) else v, v.get("ms_per_step") if isinstance(v,dict) else "")
PY
grep -i "capture\|fail" gpurun_out/r2l_bench.err | head -5
