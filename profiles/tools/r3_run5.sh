#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r3f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3f_tests.log
tail -8 gpurun_out/r3f_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3f_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r3f_smoke.log
