#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu -k "team or multi_round or round_trips" 2>&1 | tail -15 ) > gpurun_out/s5_pytest.log 2>&1
rm -f gpurun_out/timing_*.txt
for w in 2 3 4 5 6 8; do
  PGN_TIMING_DUMP=gpurun_out/timing_c2_team$w.txt PGN_TEAM=$w timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s5_bench_c2_team$w.json 2> gpurun_out/s5_bench_c2_team$w.err
done
tail -3 gpurun_out/s5_pytest.log; cut -c1-100 gpurun_out/s5_bench_*.json
