#!/bin/bash
# round 2, last 1-GPU capture at HEAD: both bench arms, the launch list of the default command, the two-leg measurement
set -x
O=gpurun_out/r2final3
mkdir -p $O
timeout 900 python bench.py > $O/bench_n1_default.json 2> $O/bench_n1_default.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_n1_reference.json 2> $O/bench_n1_reference.err
echo "reference rc=$?"
timeout 300 python tools/bench_two_legs.py > $O/two_legs.json 2> $O/two_legs.err
echo "two legs rc=$?"; cat $O/two_legs.json | head -40
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_default_headline_steps2.csv python bench.py --also "" --steps 2 --warmup 1 --no-cpu-baseline > $O/launches.log 2>&1
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "test_mala_jl" > $O/pytest_new.log 2>&1; tail -2 $O/pytest_new.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2final3/bench_n1_default.json'))
print('c3', round(d['value'],2), d['ms_per_step'], round(d['cpu_baseline']['value'],2), d['cpu_baseline']['cores'], 'e2e', round(d['e2e']['value'],2))
for k,v in d.get('also',{}).items(): print(k, round(v['value'],3), round(v['ms_per_step'],2), round(v['cpu_baseline']['value'],4))
r=json.load(open('gpurun_out/r2final3/bench_n1_reference.json'))
print('reference', round(r['value'],2), r['cpu_baseline']['cores'], {k:round(v['value'],2) for k,v in r.get('also',{}).items()})
PY
