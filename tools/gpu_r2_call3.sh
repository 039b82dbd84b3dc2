#!/bin/bash
# round 2, GPU call 3: per-replica recorders on the device (all kernel families), full GPU suite, default bench line.
set -x
O=gpurun_out/r2c3
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --durations=10 > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -25 $O/pytest_gpu.log
timeout 700 python bench.py > $O/bench_default.json 2> $O/bench_default.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c3/bench_default.json'))
print('c3', round(d['value'],2), d['ms_per_step'])
for k,v in d.get('also',{}).items(): print(k, round(v['value'],3), round(v['ms_per_step'],2))
PY
