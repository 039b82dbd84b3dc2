#!/bin/bash
# Round evidence for profiles/: bench lines of every BASELINE config, the reference arm, the launch list and
# full ncu captures of the scan kernels (run under gpurun, 1 GPU).
mkdir -p gpurun_out/prof
cd "$(dirname "$0")/.."
O=gpurun_out/prof
( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > $O/smoke.log 2>&1
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 ) > $O/pytest_gpu.log 2>&1
timeout 900 python bench.py > $O/bench_n1_c2.json 2> $O/bench_n1_c2.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_n1_c2_reference.json 2> $O/bench_n1_c2_reference.err
for c in c1 c3 c4; do timeout 600 python bench.py --config $c --no-cpu-baseline > $O/bench_n1_$c.json 2> $O/bench_n1_$c.err; done
timeout 1200 python bench.py --config c5 --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_n1_c5.json 2> $O/bench_n1_c5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 9 -c 1 -f -o $O/c2_funnel_automala python bench.py --steps 1 --warmup 1 --scans 256 --no-cpu-baseline > $O/ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 9 -c 1 -f -o $O/c3_gmm_automala python bench.py --config c3 --steps 1 --warmup 1 --scans 32 --no-cpu-baseline > $O/ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 9 -c 1 -f -o $O/c4_ising python bench.py --config c4 --steps 1 --warmup 1 --scans 64 --no-cpu-baseline > $O/ncu_c4.log 2>&1
for k in c2_funnel_automala c3_gmm_automala c4_ising; do   # keep the CSV pages, not the 30 MB reports (gpurun_out/ is capped at 64 MiB)
  ncu -i $O/$k.ncu-rep --page raw --csv > $O/${k}_ncu_raw.csv 2>/dev/null
  ncu -i $O/$k.ncu-rep --page source --csv > $O/${k}_ncu_source.csv 2>/dev/null
  rm -f $O/$k.ncu-rep
done
tail -2 $O/smoke.log $O/pytest_gpu.log; cut -c1-120 $O/bench_n1_*.json
