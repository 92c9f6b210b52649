#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_halo.py -m gpu -q --timeout 300 -p no:cacheprovider -k "1x1" -x > gpurun_out/r2q_tests.log 2>&1; echo "1x1 tests rc=$?"
tail -15 gpurun_out/r2q_tests.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu --no-also > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2q_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
print(d["roofline"]["by_entry_point_ms"])
PY
