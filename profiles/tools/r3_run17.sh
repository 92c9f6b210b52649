#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for V in 0 1; do
PC_WGRAD_TC_STEM3=$V timeout 600 python bench.py --workload train_cnn_small --steps 200 --warmup 10 --no-also --no-cpu > gpurun_out/r3s_small_$V.json 2> gpurun_out/r3s_small_$V.err; echo "small stem3_tc=$V rc=$?"
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_r2.py -m gpu -q --timeout 600 -p no:cacheprovider -k "small or phoneme_cnn or golden" > gpurun_out/r3s_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r3s_tests.log
python - <<PY
import json
for f in ["r3s_small_0","r3s_small_1"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],1), round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["roofline"]["by_entry_point_ms"].items() if "wgrad" in k})
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
