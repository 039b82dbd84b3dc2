#!/bin/bash
# round 2: the default bench line at HEAD (mixed teams on)
set -x
O=gpurun_out/r2benchlast
mkdir -p $O
timeout 600 python bench.py > $O/bench_n1_default.json 2> $O/bench_n1_default.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2benchlast/bench_n1_default.json'))
print('c3', round(d['value'],2), d['ms_per_step'], round(d['cpu_baseline']['value'],2), d['cpu_baseline']['cores'], 'e2e', round(d['e2e']['value'],2), d['roofline']['fp64']['frac'])
for k,v in d.get('also',{}).items(): print(k, round(v['value'],3), round(v['ms_per_step'],2), round(v['cpu_baseline']['value'],4))
PY
