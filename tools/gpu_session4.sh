#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/s4_pytest_full.log 2>&1
rm -f gpurun_out/timing_*.txt
for w in 1 4 8; do
  PGN_TIMING_DUMP=gpurun_out/timing_c2_team$w.txt PGN_TEAM=$w timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s4_bench_c2_team$w.json 2> gpurun_out/s4_bench_c2_team$w.err
done
for c in c1 c3 c4; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config $c > gpurun_out/s4_bench_$c.json 2> gpurun_out/s4_bench_$c.err
done
tail -3 gpurun_out/s4_pytest_full.log; cut -c1-100 gpurun_out/s4_bench_*.json
