#!/bin/bash
# round 2, 2 GPUs: sharded-ladder parity incl. the two-leg ladder whose target chains the balanced split would separate
set -x
O=gpurun_out/r2n2b
mkdir -p $O
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -k "2" > $O/pytest_multigpu.log 2>&1
echo "rc=$?" >> $O/pytest_multigpu.log; tail -8 $O/pytest_multigpu.log
