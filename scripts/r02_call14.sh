#!/bin/bash
# round 2, final 8-GPU pass: the driver's own invocation at N = 8 with the shipped defaults
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r02_final8_default.log 2>&1; echo "rc=$?"
grep '^{' gpurun_out/r02_final8_default.log | tail -1 | cut -c1-300
