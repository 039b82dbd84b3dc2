#!/bin/bash
# round 2: launch list of the default headline step at HEAD (mixed teams on)
O=gpurun_out/r2launches
mkdir -p $O
timeout 65 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_default_headline_steps2.csv python bench.py --also "" --steps 2 --warmup 1 --no-cpu-baseline > $O/launches.log 2>&1
echo "rc=$?"; grep -c scan_kernel $O/launches_bench_default_headline_steps2.csv
