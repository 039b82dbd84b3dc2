#!/bin/bash
# round 2, GPU call 6: general Compose / Mix programs, table-free Ising kernel (C4's 4096 chains on one GPU), progress-based
# spin limit; where the per-replica recorder exchange costs C2 its 4 %.
set -x
O=gpurun_out/r2c6
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -14 $O/pytest_gpu.log
B="python bench.py --no-cpu-baseline --steps 3 --warmup 3"
PGN_TIMING_DUMP=$O/timing_c2_per_replica.txt timeout 120 $B --config c2 > $O/c2_per_replica.json 2>/dev/null
PGN_TIMING_DUMP=$O/timing_c2_per_chain.txt timeout 120 $B --config c2 --recorder-order 1 > $O/c2_per_chain.json 2>/dev/null
timeout 300 $B --config c4 --scaling strong > $O/c4_strong_n1_4096.json 2> $O/c4_strong.err
PGN_ISING_LITE=1 timeout 200 $B --config c4 > $O/c4_lite_512.json 2> $O/c4_lite.err
for f in $O/c2_per_replica.json $O/c2_per_chain.json $O/c4_strong_n1_4096.json $O/c4_lite_512.json; do python -c "
import json; d=json.load(open('$f')); print('$f', round(d['value'],1), round(d['ms_per_step'],2), d['config']['n_chains'])"; done
python - <<'PY'
import numpy as np
for f in ('per_replica','per_chain'):
    a=np.loadtxt(f'gpurun_out/r2c6/timing_c2_{f}.txt'); ep=a[:,0]; last=a[ep==ep.max()]; n=last[0,2]
    e=last[:,3]/n; w=last[:,4]/n
    print(f, 'explore mean %.0f max %.0f (chain %d) | wait mean %.0f | last chain explore %.0f wait %.0f | first chain explore %.0f'%(e.mean(), e.max(), int(last[e.argmax(),1]), w.mean(), e[-1], w[-1], e[0]))
PY
