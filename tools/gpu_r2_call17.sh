#!/bin/bash
set -x
O=gpurun_out/r2c17
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x -k "unid" > $O/pytest_new.log 2>&1
echo "pytest new rc=$?"; tail -6 $O/pytest_new.log
timeout 1300 python -m pytest tests -m gpu -q --durations=3 > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -8 $O/pytest_gpu.log
