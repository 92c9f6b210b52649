#!/bin/bash
# round-2 final profile set: launch lists (cnn_deep, cnn_small) and ncu --set full captures of the conv / stem / BatchNorm / front-end kernels.
# The .ncu-rep files stay on the box (/tmp): gpurun_out/ only receives the raw-page CSV exports (64 MiB limit).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-also"
timeout 600 $CMD > gpurun_out/r2f_plain.json 2> gpurun_out/r2f_plain.err; echo "plain rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2f_launches.csv $CMD > gpurun_out/r2f_ncu.log 2>&1; echo "ncu launches rc=$?"
CMDS="python bench.py --workload train_cnn_small --steps 3 --warmup 3 --no-cpu --no-also"
timeout 600 $CMDS > gpurun_out/r2f_small_plain.json 2> gpurun_out/r2f_small_plain.err; echo "small plain rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2f_small_launches.csv $CMDS > gpurun_out/r2f_small_ncu.log 2>&1; echo "ncu small rc=$?"
CMDE="python bench.py --steps 1 --warmup 1 --no-cpu --no-also --no-graph"
timeout 1500 ncu --set full --clock-control none -k regex:"conv_halo|wgrad_halo_kernel|igemm_tc|wgrad_tc_kernel" -c 34 -o /tmp/r2f_conv_full $CMDE > gpurun_out/r2f_conv_ncu.log 2>&1; echo "ncu conv rc=$?"
ncu -i /tmp/r2f_conv_full.ncu-rep --page raw --csv > gpurun_out/r2f_conv_full_raw.csv 2>/dev/null
timeout 1500 ncu --set full --clock-control none -k regex:"stem_fwd_kernel|stem_bwd_pool|stem_gram|bn_act_bwd|bn_add_relu|bn_act_split" -c 24 -o /tmp/r2f_bn_stem_full $CMDE > gpurun_out/r2f_bn_ncu.log 2>&1; echo "ncu bn rc=$?"
ncu -i /tmp/r2f_bn_stem_full.ncu-rep --page raw --csv > gpurun_out/r2f_bn_stem_full_raw.csv 2>/dev/null
CMDF="python bench.py --workload frontend --steps 2 --warmup 1 --no-cpu"
timeout 900 ncu --set full --clock-control none -k regex:"frontend_kernel" -s 2 -c 1 -o /tmp/r2f_fe_full $CMDF > gpurun_out/r2f_fe_ncu.log 2>&1; echo "ncu fe rc=$?"
ncu -i /tmp/r2f_fe_full.ncu-rep --page raw --csv > gpurun_out/r2f_fe_full_raw.csv 2>/dev/null
ls -la gpurun_out/r2f_*; du -sh gpurun_out
