#!/bin/bash
# round 2: per-chain clocks of C3 with mixed teams (diagnostics)
O=gpurun_out/r2mixeddump
mkdir -p $O
PGN_TIMING_DUMP=$O/timing_c3_mixed.txt timeout 50 python bench.py --config c3 --no-cpu-baseline --steps 1 --warmup 1 --scans 256 > $O/c3.json 2> $O/c3.err
echo "rc=$?"; wc -l $O/timing_c3_mixed.txt
