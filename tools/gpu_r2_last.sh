#!/bin/bash
# round 2: last sanity at HEAD — smoke() and the whole GPU suite
set -x
O=gpurun_out/r2last
mkdir -p $O
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 1300 python -m pytest tests -m gpu -q --durations=3 > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -8 $O/pytest_gpu.log
