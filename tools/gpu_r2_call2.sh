#!/bin/bash
# round 2, GPU call 2: stream-ordered per-round allocations (shards on one GPU, world 3), C5 column compaction,
# register-capped C3 kernel with teams of two.
set -x
O=gpurun_out/r2c2
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --durations=10 > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
PGN_REGCAP=128 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "team_width and (toy100 or two_chains) or gmm128 or c3_gmm" > $O/pytest_regcap.log 2>&1
echo "regcap rc=$?" >> $O/pytest_regcap.log
tail -3 $O/pytest_regcap.log
B="python bench.py --config c3 --no-cpu-baseline --steps 3 --warmup 3"
timeout 120 $B > $O/c3_base.json 2> $O/c3_base.err
PGN_REGCAP=128 timeout 120 $B > $O/c3_cap128_w1.json 2> $O/c3_cap128_w1.err
PGN_REGCAP=128 PGN_TEAM=2 PGN_TIMING_DUMP=$O/timing_c3_cap128_w2.txt timeout 120 $B > $O/c3_cap128_w2.json 2> $O/c3_cap128_w2.err
timeout 300 python bench.py --config c5 --no-cpu-baseline --steps 2 --warmup 1 > $O/c5_compact.json 2> $O/c5_compact.err
for f in $O/c3_*.json $O/c5_compact.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['value'],3), round(d['ms_per_step'],2), d.get('fp64_roofline',{}).get('useful_flop_fraction'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
