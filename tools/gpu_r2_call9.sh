#!/bin/bash
# round 2, GPU call 9: explorer statistics checkpointed before the hand-shake: parity + cost on C2 / C3.
set -x
O=gpurun_out/r2c9
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
B="python bench.py --no-cpu-baseline --steps 5 --warmup 3"
timeout 120 $B --config c2 > $O/c2_per_replica.json 2>/dev/null
timeout 120 $B --config c2 --recorder-order 1 > $O/c2_per_chain.json 2>/dev/null
timeout 120 $B --config c3 > $O/c3_per_replica.json 2>/dev/null
timeout 120 $B --config c3 --recorder-order 1 > $O/c3_per_chain.json 2>/dev/null
for f in $O/c2_per_replica.json $O/c2_per_chain.json $O/c3_per_replica.json $O/c3_per_chain.json; do python -c "
import json; d=json.load(open('$f')); print('$f', round(d['value'],1), round(d['ms_per_step'],2))"; done
