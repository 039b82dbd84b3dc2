#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/timing_*.txt
for w in 4 8; do
  PGN_TIMING_DUMP=gpurun_out/timing_c2_team$w.txt PGN_TEAM=$w timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s3_bench_c2_team$w.json 2> gpurun_out/s3_bench_c2_team$w.err
done
cut -c1-120 gpurun_out/s3_bench_c2_team*.json
