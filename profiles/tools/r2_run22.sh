#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_r2.py -m gpu -q --timeout 300 -p no:cacheprovider -k "mfcc or transform or fused or frontend or dataset or noise or preemph or gain" > gpurun_out/r2v_tests.log 2>&1; echo "fe tests rc=$?"
tail -8 gpurun_out/r2v_tests.log
timeout 600 python bench.py --workload frontend --steps 10 --warmup 3 --no-cpu > gpurun_out/r2v_fe.json 2> gpurun_out/r2v_fe.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2v_fe.json"))
print(d["value"], d["ms_per_step"], d["e2e"], d["roofline"]["frac"])
PY
