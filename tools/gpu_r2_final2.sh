#!/bin/bash
# round 2, final 1-GPU evidence after the two-leg / variational change: both bench arms, the launch list of the default
# command, compute-sanitizer on the kernels new since the previous capture.
set -x
O=gpurun_out/r2final2
mkdir -p $O
timeout 900 python bench.py > $O/bench_n1_default.json 2> $O/bench_n1_default.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_n1_reference.json 2> $O/bench_n1_reference.err
echo "reference rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_default_headline_steps2.csv python bench.py --also "" --steps 2 --warmup 1 --no-cpu-baseline > $O/launches.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shards.py -m gpu -q -x -k "two_legs_gmm6_automala_gaussian or two_legs_funnel8_slice_gaussian or two_legs_ising5 or two_legs_funnel_slice_n10 or integrator and toy_mvn or variational_entry" > $O/compute_sanitizer_memcheck.log 2>&1
echo "sanitizer rc=$?" >> $O/compute_sanitizer_memcheck.log; tail -6 $O/compute_sanitizer_memcheck.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2final2/bench_n1_default.json'))
print('c3', round(d['value'],2), d['ms_per_step'], round(d['cpu_baseline']['value'],2), d['cpu_baseline']['cores'], 'e2e', round(d['e2e']['value'],2))
for k,v in d.get('also',{}).items(): print(k, round(v['value'],3), round(v['ms_per_step'],2), round(v['cpu_baseline']['value'],4))
r=json.load(open('gpurun_out/r2final2/bench_n1_reference.json'))
print('reference', round(r['value'],2), r['cpu_baseline']['cores'], {k:round(v['value'],2) for k,v in r.get('also',{}).items()})
PY
