#!/bin/bash
# round 2, GPU call 4: MIXED target (Bool / Integer slice coordinates), recorder-order cost on C2, ncu evidence for C3.
set -x
O=gpurun_out/r2c4
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --durations=5 > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -8 $O/pytest_gpu.log
B="python bench.py --no-cpu-baseline --steps 5 --warmup 3"
timeout 120 $B --config c2 > $O/c2_per_replica.json 2> $O/c2_per_replica.err
timeout 120 $B --config c2 --recorder-order 1 > $O/c2_per_chain.json 2> $O/c2_per_chain.err
timeout 120 $B --config c3 --recorder-order 1 > $O/c3_per_chain.json 2> $O/c3_per_chain.err
timeout 120 $B --config c3 > $O/c3_per_replica.json 2> $O/c3_per_replica.err
for f in $O/c[23]_per_*.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', round(d['value'],1), round(d['ms_per_step'],2))"; done
# ncu: launch list of the default command (headline only), then one full capture of the C3 kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_c3_steps2.csv python bench.py --also "" --steps 2 --warmup 1 --no-cpu-baseline > $O/launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 9 -c 1 -f -o $O/c3_gmm_automala python bench.py --config c3 --steps 1 --warmup 1 --scans 32 --no-cpu-baseline > $O/ncu_c3.log 2>&1
ncu -i $O/c3_gmm_automala.ncu-rep --page raw --csv > $O/c3_gmm_automala_ncu_raw.csv 2>/dev/null
ncu -i $O/c3_gmm_automala.ncu-rep --page source --csv > $O/c3_gmm_automala_ncu_source.csv 2>/dev/null
rm -f $O/c3_gmm_automala.ncu-rep
ls -la $O
