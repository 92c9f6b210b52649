#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_gpu_halo.py tests/test_gpu_parity.py tests/test_gpu_r2.py tests/test_gpu_tc.py -m gpu -q --timeout 300 -p no:cacheprovider -k "halo_wgrad or nets or bench_shape or training_step or graph or trainer or small" > gpurun_out/r2aa_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r2aa_tests.log
timeout 600 python bench.py --workload train_cnn_small --steps 100 --warmup 10 --no-cpu > gpurun_out/r2aa_small.json 2> gpurun_out/r2aa_small.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2aa_small.json"))
print(d["value"], d["ms_per_step"], d["e2e"])
print(d["roofline"].get("by_entry_point_ms"))
PY
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-also > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2aa_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
PY
