#!/bin/bash
# round 2: two legs (StabilizedPT) + GaussianReference + the device integrator entry point — targeted parity, then a
# C3 / C2 bench to confirm the plain kernels did not move
set -x
O=gpurun_out/r2c13
mkdir -p $O
timeout 700 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shards.py tests/test_golden.py -m gpu -q -x \
  -k "two_legs or one_leg or variational or integrator or involution or refuse" > $O/pytest_new.log 2>&1
echo "pytest new rc=$?"; tail -15 $O/pytest_new.log
timeout 150 python bench.py --config c3 --no-cpu-baseline --steps 3 --warmup 3 > $O/c3.json 2> $O/c3.err
timeout 150 python bench.py --config c2 --no-cpu-baseline --steps 3 --warmup 3 > $O/c2.json 2> $O/c2.err
for f in $O/c3.json $O/c2.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d['value'],1), d['ms_per_step'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
