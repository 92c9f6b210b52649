#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/r3v_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r3v_tests.log
timeout 600 python bench.py --steps 100 --warmup 5 --no-also --no-cpu > gpurun_out/r3v_deep.json 2> gpurun_out/r3v_deep.err; echo "deep rc=$?"
timeout 600 python bench.py --steps 100 --warmup 5 --no-also --no-cpu > gpurun_out/r3v_deep2.json 2> gpurun_out/r3v_deep2.err; echo "deep2 rc=$?"
python - <<PY
import json
for f in ["r3v_deep","r3v_deep2"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
