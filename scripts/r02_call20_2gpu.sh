#!/bin/bash
# round 2, after the cache-hint change: the driver's own invocation at N = 2 (multirank bit-compare, cfg2 weak, cfg3 strong)
mkdir -p gpurun_out
( time timeout 95 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29720 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/r02b_final2_default.log 2>&1; echo "rc=$?"
grep '^{' gpurun_out/r02b_final2_default.log | tail -1 | cut -c1-300
