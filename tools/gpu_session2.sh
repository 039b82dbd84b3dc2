#!/bin/bash
mkdir -p gpurun_out
PGN_TEAM=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 9 -c 1 -o gpurun_out/r01_c2_team4 python bench.py --steps 1 --warmup 1 --scans 256 --no-cpu-baseline > gpurun_out/s2_ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 9 -c 1 -o gpurun_out/r01_c4_ising python bench.py --config c4 --steps 1 --warmup 1 --scans 64 --no-cpu-baseline > gpurun_out/s2_ncu_c4.log 2>&1
tail -3 gpurun_out/s2_ncu_c2.log gpurun_out/s2_ncu_c4.log; ls -la gpurun_out
