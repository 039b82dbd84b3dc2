#!/bin/bash
# one gpurun call: parity tests, then bench lines for the team widths and the other configs
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu -k "team or ising or multi_round or round_trips" 2>&1 | tail -25 ) > gpurun_out/s1_pytest_new.log 2>&1
for w in 1 2 4 6 8; do
  PGN_TEAM=$w timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s1_bench_c2_team$w.json 2> gpurun_out/s1_bench_c2_team$w.err
done
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config c4 > gpurun_out/s1_bench_c4.json 2> gpurun_out/s1_bench_c4.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config c3 > gpurun_out/s1_bench_c3.json 2> gpurun_out/s1_bench_c3.err
( timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 ) > gpurun_out/s1_pytest_full.log 2>&1
tail -5 gpurun_out/s1_pytest_new.log; cat gpurun_out/s1_bench_c2_team*.json | cut -c1-200; tail -3 gpurun_out/s1_pytest_full.log
