#!/bin/bash
# round 2, final 1-GPU evidence for profiles/: smoke, the full GPU suite, both bench arms, the launch list of the default
# command, one full ncu capture of the tensor-core GEMM, compute-sanitizer on the kernels new in this round.
set -x
O=gpurun_out/r2final
mkdir -p $O
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > $O/smoke.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench_n1_default.json 2> $O/bench_n1_default.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_n1_reference.json 2> $O/bench_n1_reference.err
echo "reference rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_default_headline_steps2.csv python bench.py --also "" --steps 2 --warmup 1 --no-cpu-baseline > $O/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dgemm_km_dmma -s 40 -c 1 -f -o $O/c5_dgemm_dmma python bench.py --config c5 --steps 1 --warmup 0 --burn-rounds 1 --no-cpu-baseline > $O/ncu_c5.log 2>&1
ncu -i $O/c5_dgemm_dmma.ncu-rep --page raw --csv > $O/c5_dgemm_dmma_ncu_raw.csv 2>/dev/null
rm -f $O/c5_dgemm_dmma.ncu-rep
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mixed_bool_int_float_slice or toy5_compose_mala_slice_automala or ising_table_free_kernel_parity and ising5 or test_multi_round_parity and (c1_toy_slice or toy10_automala or logreg24_automala or ising5)" > $O/compute_sanitizer_memcheck.log 2>&1
echo "sanitizer rc=$?" >> $O/compute_sanitizer_memcheck.log; tail -6 $O/compute_sanitizer_memcheck.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2final/bench_n1_default.json'))
print('c3', round(d['value'],2), d['ms_per_step'], round(d['cpu_baseline']['value'],2), d['cpu_baseline']['cores'])
for k,v in d.get('also',{}).items(): print(k, round(v['value'],3), round(v['ms_per_step'],2), round(v['cpu_baseline']['value'],4))
r=json.load(open('gpurun_out/r2final/bench_n1_reference.json'))
print('reference', round(r['value'],2), r['cpu_baseline']['cores'], {k:round(v['value'],2) for k,v in r.get('also',{}).items()})
PY
