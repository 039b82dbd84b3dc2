#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/s16_pytest_full.log 2>&1
for c in c2 c3; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config $c > gpurun_out/s16_bench_$c.json 2>gpurun_out/s16_bench_$c.err
done
tail -4 gpurun_out/s16_pytest_full.log; cut -c1-100 gpurun_out/s16_bench_*.json
