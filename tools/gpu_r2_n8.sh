#!/bin/bash
# round 2, 8-GPU call: sharded-ladder parity at world 8 and the default bench line at BASELINE's multi-GPU shapes
# (C4: 4096 chains, C5: 2048 chains; C3 / C2 weak-scaled), strong scaling of the C3 ladder.
set -x
O=gpurun_out/r2n8
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -k "8" > $O/pytest_multigpu.log 2>&1
echo "rc=$?" >> $O/pytest_multigpu.log; tail -5 $O/pytest_multigpu.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $T --master-port 29631 bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err
echo "bench n8 rc=$?"
timeout 200 $T --master-port 29632 bench.py --gpus 8 --config c3 --scaling strong --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_n8_c3_strong.json 2> $O/bench_n8_c3_strong.err
for f in $O/bench_n8.json $O/bench_n8_c3_strong.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['value'],2), d['config']['n_chains'], d['scaling'], {k:(round(v['value'],3), v['config']['n_chains']) for k,v in d.get('also',{}).items()})
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
tail -3 $O/bench_n8.err
