#!/bin/bash
# round 2, GPU call 10: mixture eval_grad with two passes of four modes (bit-identical, fewer live accumulators);
# the same kernel at 144 registers with teams of two on C3.
set -x
O=gpurun_out/r2c10
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "gmm or c3 or golden or team_width" > $O/pytest_gmm.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gmm.log; tail -4 $O/pytest_gmm.log
PGN_REGCAP=144 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gmm128 or c3_gmm or (team_width and toy100)" > $O/pytest_regcap.log 2>&1
echo "regcap rc=$?" >> $O/pytest_regcap.log; tail -3 $O/pytest_regcap.log
B="python bench.py --config c3 --no-cpu-baseline --steps 4 --warmup 3"
timeout 120 $B > $O/c3_w1.json 2>/dev/null
PGN_REGCAP=144 PGN_TEAM=2 PGN_TIMING_DUMP=$O/timing_c3_cap144_w2.txt timeout 120 $B > $O/c3_cap144_w2.json 2>/dev/null
PGN_REGCAP=144 PGN_TEAM=1 timeout 120 $B > $O/c3_cap144_w1.json 2>/dev/null
timeout 120 $B > $O/c3_w1_again.json 2>/dev/null
for f in $O/c3_*.json; do python -c "
import json; d=json.load(open('$f')); print('$f', round(d['value'],1), round(d['ms_per_step'],2))"; done
python - <<'PY'
import numpy as np
a=np.loadtxt('gpurun_out/r2c10/timing_c3_cap144_w2.txt'); ep=a[:,0]; last=a[ep==ep.max()]; n=last[0,2]
for nm,col in (('explore',3),('wait',4),('points',5),('trial',6),('barrier',7),('decide',8)):
    v=last[:,col]/n; print(nm, 'mean %.0f max %.0f'%(v.mean(), v.max()))
PY
