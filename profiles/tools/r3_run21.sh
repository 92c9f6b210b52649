#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/r3ae_tests.log 2>&1; echo "tests rc=$?"
grep -v "^E    \|^    " gpurun_out/r3ae_tests.log | tail -8
timeout 600 python bench.py --workload train_cnn_small --steps 200 --warmup 10 --no-also --no-cpu > gpurun_out/r3ae_small.json 2> gpurun_out/r3ae_small.err; echo "small rc=$?"
python - <<PY
import json
for f in ["r3ae_small"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), {k:round(v,4) for k,v in d["roofline"]["by_entry_point_ms"].items() if "bn_act_bwd" in k})
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
