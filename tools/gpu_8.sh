#!/bin/bash
# 8-GPU evidence: sharded-ladder parity at world 2/4/8, then the weak-scaling bench lines
mkdir -p gpurun_out/n8
O=gpurun_out/n8
( timeout 900 python -m pytest tests/test_multigpu.py -x -q -m gpu 2>&1 | tail -6 ) > $O/pytest_multigpu_8gpus.log 2>&1
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/bench_n${n}_c2.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 8 --config c4 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/bench_n8_c4.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29660 bench.py --gpus 8 --config c5 --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | tail -1 > $O/bench_n8_c5.json
tail -3 $O/pytest_multigpu_8gpus.log; cut -c1-110 $O/bench_*.json
