#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_halo.py -m gpu -q --timeout 300 -p no:cacheprovider -k "wgrad" > gpurun_out/r3n_halo.log 2>&1; echo "wgrad tests rc=$?"
tail -3 gpurun_out/r3n_halo.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r3n_tests.log 2>&1; echo "all tests rc=$?"
grep -v "^E    \|^    " gpurun_out/r3n_tests.log | tail -25
for V in 1 0; do
PC_SMALL_C32=$V timeout 600 python bench.py --workload train_cnn_small --steps 200 --warmup 10 --no-also --no-cpu > gpurun_out/r3n_small_$V.json 2> gpurun_out/r3n_small_$V.err; echo "small $V rc=$?"
done
python - <<PY
import json
for f in ["r3n_small_1","r3n_small_0"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1))
        print({k:round(v,4) for k,v in d["roofline"]["by_entry_point_ms"].items() if "conv" in k})
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
