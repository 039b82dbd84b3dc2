#!/bin/bash
# round 2: the unidentifiable-product target (the reference's own test target) + the whole GPU suite again
set -x
O=gpurun_out/r2c16
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x -k "unid or dlp or device_math or numerics" > $O/pytest_new.log 2>&1
echo "pytest new rc=$?"; tail -6 $O/pytest_new.log
timeout 1300 python -m pytest tests -m gpu -q --durations=5 > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -12 $O/pytest_gpu.log
