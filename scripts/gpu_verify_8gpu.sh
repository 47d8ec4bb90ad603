#!/bin/bash
# 8-GPU pass (not run in round 1: no 8-GPU slot was taken): multi-rank parity with all 8 ranks
# (octants: face, edge and corner peers), weak-scaled cavity = the 512^3 mesh of north_star, cfg3
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_multirank.py -m gpu -x -q ) > gpurun_out/pytest_multi8.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_multi8.log
# the bench's own 8-octant mesh at level 5, peer-memory exchange, 40 steps, bit-compared with the oracle
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 tests/parity_multi.py --mode gpu --layout d3q19 --relaxation trt --kind cavity --level 5 --steps 40 --octants 8 --p2p > gpurun_out/parity_8oct.log 2>&1
grep -c "ndiff=0" gpurun_out/parity_8oct.log
run() { label=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 8 --steps 300 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/bench_8gpu_$label.log 2>&1; }
run p2p
run nccl --no-p2p --no-e2e
run cfg3_p2p --workload cfg3 --no-e2e --steps 100
tail -3 gpurun_out/pytest_multi8.log
for f in gpurun_out/bench_8gpu_*.log; do echo $f; tail -1 $f | cut -c1-200; done
