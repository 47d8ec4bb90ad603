#!/bin/bash
# 2-GPU pass for the multi-level path: parity (2 and 3 levels) + one single-level case, cfg4 bench
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_multirank.py -m gpu -x -q -k "multilevel or trt19-cavity-2oct-p2p" ) > gpurun_out/pytest_multi_ml.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_multi_ml.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 2 --workload cfg4 --steps 100 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg4_2gpu.log 2>&1
# NOT YET RUN ON A GPU (added after the round's GPU budget was spent): shared ghosts delegated to one
# rank and exchanged through the FromCoarser / FromFiner buffers (state + auxField of ghostFromFiner);
# the oracle's multi-rank driver is bit-identical with it (tests/test_multilevel.py, CPU)
for lv in 2 3; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29652 tests/parity_multi.py --mode gpu-ml --layout d3q19 --relaxation bgk --levels $lv --steps 6 --ghost-exchange > gpurun_out/parity_ghost_exchange_$lv.log 2>&1
  grep -h "ghosts received\|multilevel" gpurun_out/parity_ghost_exchange_$lv.log | tail -4
done
tail -6 gpurun_out/pytest_multi_ml.log; tail -1 gpurun_out/bench_cfg4_2gpu.log | cut -c1-300
