#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider > gpurun_out/r3ai_tests.log 2>&1; echo "tests rc=$?"
tail -2 gpurun_out/r3ai_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
