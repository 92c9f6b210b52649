#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_r2.py tests/test_gpu_tc.py -m gpu -q --timeout 300 -p no:cacheprovider -k "nets or bench_shape or training_step or bn" -x > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2x_tests.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu --no-also > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2x_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
print(d["roofline"]["by_entry_point_ms"])
PY
