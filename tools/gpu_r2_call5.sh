#!/bin/bash
# round 2, GPU call 5: tensor-core GEMM with 128 x 64 tiles (two blocks per SM) vs 128 x 128, LOGREG parity, C5 bench.
set -x
O=gpurun_out/r2c5
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -k "logreg or c5 or golden or online or recorder" > $O/pytest_logreg.log 2>&1
echo "pytest rc=$?" >> $O/pytest_logreg.log; tail -4 $O/pytest_logreg.log
B="python bench.py --config c5 --no-cpu-baseline --steps 2 --warmup 1"
timeout 300 $B > $O/c5_bn64.json 2> $O/c5_bn64.err
PGN_DMMA_BN=128 timeout 300 $B > $O/c5_bn128.json 2> $O/c5_bn128.err
for f in $O/c5_bn64.json $O/c5_bn128.json; do python -c "
import json; d=json.load(open('$f')); r=d['fp64_roofline']; print('$f', round(d['value'],4), round(d['ms_per_step'],1), round(r['achieved_tflops'],2), round(r['frac'],3), round(r['useful_flop_fraction'],3), round(r['gemm_share_of_kernel_time'],3))"; done
