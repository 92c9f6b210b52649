#!/bin/bash
# launch list of one cnn_deep bench run (per-launch gpu time) after a plain run of the same command
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-also"
timeout 600 $CMD > gpurun_out/r2m_plain.json 2> gpurun_out/r2m_plain.err; echo "plain rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2m_launches.csv $CMD > gpurun_out/r2m_ncu.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r2m_launches.csv
