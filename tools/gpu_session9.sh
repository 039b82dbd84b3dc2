#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu -k "team or multi_round or round_trips" 2>&1 | tail -15 ) > gpurun_out/s9_pytest.log 2>&1
for c in c2 c3; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config $c > gpurun_out/s9_bench_$c.json 2> gpurun_out/s9_bench_$c.err
done
for w in 4 5 8; do
PGN_TEAM=$w timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s9_bench_c2_team$w.json 2>/dev/null
done
tail -3 gpurun_out/s9_pytest.log; cut -c1-100 gpurun_out/s9_bench_*.json
