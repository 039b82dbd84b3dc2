#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu -k "ising" 2>&1 | tail -15 ) > gpurun_out/s14_pytest.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config c4 > gpurun_out/s14_bench_c4.json 2>gpurun_out/s14_bench_c4.err
tail -3 gpurun_out/s14_pytest.log; cut -c1-100 gpurun_out/s14_bench_*.json
