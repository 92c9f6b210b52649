#!/bin/bash
# final-state launch lists (cnn_deep, cnn_small) and full captures of the new attention-pool kernels
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-also"
timeout 600 $CMD > gpurun_out/r3j_plain.json 2> gpurun_out/r3j_plain.err; echo "plain rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r3j_launches.csv $CMD > gpurun_out/r3j_ncu.log 2>&1; echo "ncu launches rc=$?"
CMDS="python bench.py --workload train_cnn_small --steps 3 --warmup 3 --no-cpu --no-also"
timeout 600 $CMDS > gpurun_out/r3j_small_plain.json 2> gpurun_out/r3j_small_plain.err; echo "small plain rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r3j_small_launches.csv $CMDS > gpurun_out/r3j_small_ncu.log 2>&1; echo "ncu small rc=$?"
CMDE="python bench.py --workload train_cnn_small --steps 1 --warmup 1 --no-cpu --no-also --no-graph"
timeout 900 ncu --set full --clock-control none -k regex:"attn_pool" -c 4 -o /tmp/r3j_attn_full $CMDE > gpurun_out/r3j_attn_ncu.log 2>&1; echo "ncu attn rc=$?"
ncu -i /tmp/r3j_attn_full.ncu-rep --page raw --csv > gpurun_out/r3j_attn_full_raw.csv 2>/dev/null
ls -la gpurun_out/r3j_*
