#!/bin/bash
# round 2, GPU call 1: full GPU test-suite on HEAD (TU split, new shard / resume / C5 full-shape tests), the default
# bench line, and per-chain timing dumps of the C3 kernel.
set -x
mkdir -p gpurun_out/r2c1
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r2c1/gpu.txt
timeout 1100 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2c1/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c1/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2c1/bench_default.json 2> gpurun_out/r2c1/bench_default.err
echo "bench rc=$?"
PGN_TIMING_DUMP=gpurun_out/r2c1/timing_c3.txt timeout 200 python bench.py --config c3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2c1/bench_c3_dump.json 2>&1
tail -5 gpurun_out/r2c1/pytest_gpu.log
