#!/bin/bash
mkdir -p gpurun_out/prof
O=gpurun_out/prof
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 9 -c 1 -f -o $O/c4_ising python bench.py --config c4 --steps 1 --warmup 1 --scans 64 --no-cpu-baseline > $O/ncu_c4.log 2>&1
ncu -i $O/c4_ising.ncu-rep --page raw --csv > $O/c4_ising_ncu_raw.csv 2>/dev/null
ncu -i $O/c4_ising.ncu-rep --page source --csv > $O/c4_ising_ncu_source.csv 2>/dev/null
rm -f $O/c4_ising.ncu-rep
