#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_r2.py tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r3g_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r3g_tests.log
timeout 600 python bench.py --workload train_cnn_small --steps 100 --warmup 10 --no-also --no-cpu > gpurun_out/r3g_small.json 2> gpurun_out/r3g_small.err; echo "small rc=$?"
timeout 600 python bench.py --steps 50 --warmup 5 --no-also --no-cpu > gpurun_out/r3g_deep.json 2> gpurun_out/r3g_deep.err; echo "deep rc=$?"
python - <<PY
import json
for f in ["r3g_small","r3g_deep"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],1), round(d["ms_per_step"],4), round(d["e2e"]["value"],1))
        for k,v in sorted(d["roofline"].get("entries",{}).items(), key=lambda kv:-kv[1].get("ms",0))[:30]:
            print("   ", k, round(v["ms"],4), v.get("calls"))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
