#!/bin/bash
set -x
O=gpurun_out/r2c18
mkdir -p $O
timeout 1300 python -m pytest tests -m gpu -q --durations=3 > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -8 $O/pytest_gpu.log
