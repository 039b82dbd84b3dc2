#!/bin/bash
# round 2, 2-GPU call: sharded-ladder parity (IPC mailboxes over NVLink) with checked_round, and the bench contract
# under torchrun (both arms, weak and strong scaling).
set -x
O=gpurun_out/r2n2
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q > $O/pytest_multigpu.log 2>&1
echo "rc=$?" >> $O/pytest_multigpu.log; tail -5 $O/pytest_multigpu.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $T --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
echo "bench n2 rc=$?"
timeout 600 $T --master-port 29612 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > $O/bench_n2_reference.json 2> $O/bench_n2_reference.err
echo "ref n2 rc=$?"
timeout 300 $T --master-port 29613 bench.py --gpus 2 --config c2 --scaling strong --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_n2_c2_strong.json 2> $O/bench_n2_c2_strong.err
timeout 300 $T --master-port 29614 bench.py --gpus 2 --config c3 --scaling strong --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_n2_c3_strong.json 2> $O/bench_n2_c3_strong.err
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d.get('impl','b200'), round(d['value'],2), d['config']['n_chains'], d['scaling'], d.get('cpu_baseline',{}).get('cores'), {k:round(v['value'],3) for k,v in d.get('also',{}).items()})
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
tail -3 $O/*.err
