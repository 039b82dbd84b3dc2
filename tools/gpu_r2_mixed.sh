#!/bin/bash
# round 2: "mixed teams" launch of single-warp autoMALA ladders — parity (forced on small ladders, and C3 at full width
# against the oracle) and the C3 bench with and without it
set -x
O=gpurun_out/r2mixed
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mixed_teams" --durations=3 > $O/pytest_mixed.log 2>&1
echo "pytest mixed rc=$?"; tail -8 $O/pytest_mixed.log
PGN_MIXED_TEAMS=1 timeout 150 python bench.py --config c3 --no-cpu-baseline --steps 3 --warmup 3 > $O/c3_mixed.json 2> $O/c3_mixed.err
PGN_MIXED_TEAMS=0 timeout 150 python bench.py --config c3 --no-cpu-baseline --steps 3 --warmup 3 > $O/c3_plain.json 2> $O/c3_plain.err
for f in $O/c3_mixed.json $O/c3_plain.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d['value'],1), d['ms_per_step'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
tail -3 $O/c3_mixed.err
