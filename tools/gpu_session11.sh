#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu -k "ising or team or multi_round" 2>&1 | tail -15 ) > gpurun_out/s12_pytest.log 2>&1
rm -f gpurun_out/timing_*.txt
PGN_TIMING_DUMP=gpurun_out/timing_c2_team6.txt timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s12_bench_c2.json 2>/dev/null
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config c4 > gpurun_out/s12_bench_c4.json 2>/dev/null
tail -3 gpurun_out/s12_pytest.log; cut -c1-100 gpurun_out/s12_bench_*.json
