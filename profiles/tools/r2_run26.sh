#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_halo.py tests/test_gpu_parity.py tests/test_gpu_r2.py tests/test_gpu_tc.py -m gpu -q --timeout 300 -p no:cacheprovider -k "halo_wgrad or nets or bench_shape or training_step or graph" -x > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2z_tests.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu --no-also > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2z_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
print(d["roofline"]["by_entry_point_ms"])
PY
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-also"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2z_launches.csv $CMD > gpurun_out/r2z_ncu.log 2>&1; echo "ncu rc=$?"
