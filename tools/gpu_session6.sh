#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 9 -c 1 -o gpurun_out/r01_c2_team6 -f python bench.py --steps 1 --warmup 1 --scans 256 --no-cpu-baseline > gpurun_out/s6_ncu_c2.log 2>&1
ls -la gpurun_out | head
