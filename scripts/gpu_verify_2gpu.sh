#!/bin/bash
# 2-GPU pass: multi-rank parity incl. the push fused into the sweep, fused vs separate push kernel
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_multirank.py -m gpu -x -q ) > gpurun_out/pytest_multi.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
run() { label=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 2 --steps 300 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/bench_2gpu_$label.log 2>&1; }
run fused --fused-push --no-e2e
run unfused --no-e2e
run fused2 --fused-push
tail -3 gpurun_out/pytest_multi.log
for f in gpurun_out/bench_2gpu_fused.log gpurun_out/bench_2gpu_unfused.log gpurun_out/bench_2gpu_fused2.log; do echo $f; tail -1 $f | cut -c1-200; done
