#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/s8_pytest_full.log 2>&1
for c in c2 c3 c4 c1; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config $c > gpurun_out/s8_bench_$c.json 2> gpurun_out/s8_bench_$c.err
done
PGN_TEAM=4 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s8_bench_c2_team4.json 2>/dev/null
PGN_TEAM=8 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s8_bench_c2_team8.json 2>/dev/null
tail -3 gpurun_out/s8_pytest_full.log; cut -c1-100 gpurun_out/s8_bench_*.json
