#!/bin/bash
# round 2, 4-GPU call: sharded-ladder parity at world 2 and 4 (middle ranks have two neighbours), the default bench line
# under torchrun, strong scaling of the C3 ladder, and per-chain wait clocks of the C2 ladder (boundary vs interior chains).
set -x
O=gpurun_out/r2n4
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q > $O/pytest_multigpu.log 2>&1
echo "rc=$?" >> $O/pytest_multigpu.log; tail -5 $O/pytest_multigpu.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 900 $T --master-port 29621 bench.py --gpus 4 --steps 5 --warmup 3 > $O/bench_n4.json 2> $O/bench_n4.err
echo "bench n4 rc=$?"
timeout 300 $T --master-port 29622 bench.py --gpus 4 --config c3 --scaling strong --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_n4_c3_strong.json 2> $O/bench_n4_c3_strong.err
timeout 300 $T --master-port 29623 bench.py --gpus 4 --config c4 --scaling strong --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_n4_c4_strong.json 2> $O/bench_n4_c4_strong.err
PGN_TIMING_DUMP=$O/timing_c2_n4.txt timeout 300 $T --master-port 29624 bench.py --gpus 4 --config c2 --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_n4_c2_dump.json 2> $O/bench_n4_c2_dump.err
for f in $O/bench_n4.json $O/bench_n4_c3_strong.json $O/bench_n4_c4_strong.json $O/bench_n4_c2_dump.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['value'],2), d['config']['n_chains'], d['scaling'], {k:round(v['value'],3) for k,v in d.get('also',{}).items()})
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
