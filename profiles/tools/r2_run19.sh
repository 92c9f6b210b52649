#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_halo.py tests/test_gpu_r2.py -m gpu -q --timeout 300 -p no:cacheprovider -k "halo or stem" -x > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2s_tests.log
timeout 300 python profiles/tools/stem_bench.py 2>&1 | tee gpurun_out/r2s_stem_bench.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu --no-also > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2s_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
print(d["roofline"]["by_entry_point_ms"])
PY
