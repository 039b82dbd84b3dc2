#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu -k "team or multi_round or round_trips" 2>&1 | tail -15 ) > gpurun_out/s10_pytest.log 2>&1
rm -f gpurun_out/timing_*.txt
for w in 4 5 6; do
PGN_TIMING_DUMP=gpurun_out/timing_c2_team$w.txt PGN_TEAM=$w timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s10_bench_c2_team$w.json 2>/dev/null
done
tail -3 gpurun_out/s10_pytest.log; cut -c1-100 gpurun_out/s10_bench_*.json
