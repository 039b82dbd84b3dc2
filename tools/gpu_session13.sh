#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/s13_pytest_full.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --config c5 > gpurun_out/s13_bench_c5.json 2> gpurun_out/s13_bench_c5.err
for c in c1 c3; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config $c > gpurun_out/s13_bench_$c.json 2>/dev/null
done
tail -3 gpurun_out/s13_pytest_full.log; cut -c1-100 gpurun_out/s13_bench_*.json
