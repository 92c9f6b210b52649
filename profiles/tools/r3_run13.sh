#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMDS="python bench.py --workload train_cnn_small --steps 3 --warmup 3 --no-cpu --no-also"
timeout 600 $CMDS > gpurun_out/r3o_small_plain.json 2> gpurun_out/r3o_small_plain.err; echo "small plain rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r3o_small_launches.csv $CMDS > gpurun_out/r3o_small_ncu.log 2>&1; echo "ncu small rc=$?"
CMDE="python bench.py --workload train_cnn_small --steps 1 --warmup 1 --no-cpu --no-also --no-graph"
timeout 900 ncu --set full --clock-control none -k regex:"bn_act_bwd|bn_act_fwd|conv_halo_res_kernel|wgrad_halo_kernel" -c 30 -o /tmp/r3o_small_full $CMDE > gpurun_out/r3o_small_full_ncu.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/r3o_small_full.ncu-rep --page raw --csv > gpurun_out/r3o_small_full_raw.csv 2>/dev/null
ls -la gpurun_out/r3o_*
