#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_halo.py -m gpu -q --timeout 600 -p no:cacheprovider -k "fused_bn_reduce" > gpurun_out/r3i_halo.log 2>&1; echo "halo tests rc=$?"
tail -25 gpurun_out/r3i_halo.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_r2.py tests/test_gpu_tc.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r3i_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r3i_tests.log
for V in 1 0; do
PC_DGRAD_BNRED=$V timeout 600 python bench.py --steps 100 --warmup 5 --no-also --no-cpu > gpurun_out/r3i_deep_$V.json 2> gpurun_out/r3i_deep_$V.err; echo "deep $V rc=$?"
done
python - <<PY
import json
for f in ["r3i_deep_1","r3i_deep_0"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1))
        print({k:round(v,4) for k,v in d["roofline"]["by_entry_point_ms"].items() if "bn_act" in k or "dgrad" in k})
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
