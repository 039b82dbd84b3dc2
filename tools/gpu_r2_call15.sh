#!/bin/bash
set -x
O=gpurun_out/r2c15
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "two_legs" --durations=5 > $O/pytest_new.log 2>&1
echo "pytest new rc=$?"; tail -12 $O/pytest_new.log
