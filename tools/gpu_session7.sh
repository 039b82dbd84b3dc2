#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 9 -c 1 -o gpurun_out/r01_c3_gmm -f python bench.py --config c3 --steps 1 --warmup 1 --scans 32 --no-cpu-baseline > gpurun_out/s7_ncu_c3.log 2>&1
tail -2 gpurun_out/s7_ncu_c3.log
