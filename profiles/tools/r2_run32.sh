#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_halo.py -m gpu -q --timeout 120 -p no:cacheprovider -k "stride2" > gpurun_out/r2ae_tests.log 2>&1; echo "s2 tests rc=$?"
tail -12 gpurun_out/r2ae_tests.log | cut -c1-200
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu --no-also > gpurun_out/r2ae_bench.json 2> gpurun_out/r2ae_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2ae_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
print(d["roofline"]["by_entry_point_ms"])
PY
PC_HALO_S2=0 timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu --no-also > gpurun_out/r2ae_bench_off.json 2> gpurun_out/r2ae_bench_off.err
PC_HALO_ALL=1 timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu --no-also > gpurun_out/r2ae_bench_all.json 2> gpurun_out/r2ae_bench_all.err
python - <<'PY'
import json
for f in ("off","all"):
    d=json.load(open(f"gpurun_out/r2ae_bench_{f}.json"))
    print(f, d["value"], d["ms_per_step"], d["roofline"]["by_entry_point_ms"]["pc_conv_dgrad"], d["roofline"]["by_entry_point_ms"]["pc_conv_fwd"])
PY
