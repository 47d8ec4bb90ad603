#!/bin/bash
# 4-GPU pass: multi-rank parity (4 ranks: face, edge peers), weak-scaled cavity, cfg3
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_multirank.py -m gpu -x -q ) > gpurun_out/pytest_multi4.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_multi4.log
run() { label=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 4 --steps 300 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/bench_4gpu_$label.log 2>&1; }
run p2p
run nccl --no-p2p --no-e2e
run cfg3_p2p --workload cfg3 --no-e2e --steps 100
tail -3 gpurun_out/pytest_multi4.log
for f in gpurun_out/bench_4gpu_*.log; do echo $f; tail -1 $f | cut -c1-200; done
