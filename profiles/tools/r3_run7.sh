#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_r2.py -m gpu -q --timeout 600 -p no:cacheprovider -k "trainer or graph or epoch" > gpurun_out/r3h_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r3h_tests.log
timeout 600 python bench.py --steps 100 --warmup 5 --no-also --no-cpu > gpurun_out/r3h_deep.json 2> gpurun_out/r3h_deep.err; echo "deep rc=$?"
timeout 600 python bench.py --workload train_cnn_small --steps 200 --warmup 10 --no-also --no-cpu > gpurun_out/r3h_small.json 2> gpurun_out/r3h_small.err; echo "small rc=$?"
python - <<PY
import json
for f in ["r3h_deep","r3h_small"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
