#!/bin/bash
# round 2: the whole GPU suite after the two-leg / variational / shard-layout changes
set -x
O=gpurun_out/r2c14
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shards.py -m gpu -q -x -k "two_legs or refuse" > $O/pytest_new.log 2>&1
echo "pytest new rc=$?"; tail -5 $O/pytest_new.log
timeout 1300 python -m pytest tests -m gpu -q --durations=8 > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -14 $O/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
