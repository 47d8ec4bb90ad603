#!/bin/bash
# round 2, 2 GPUs: multi-rank parity (all exchange paths, multi-level p2p, restart, ghost exchange),
# the driver's own bench invocation at N = 2, cfg4 strong scaling p2p vs NCCL
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_multirank.py -m gpu -q --durations=5 ) > gpurun_out/r02_pytest_multi2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_multi2.log
tail -15 gpurun_out/r02_pytest_multi2.log
run() { label=$1; shift; ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 2 "$@" ) > gpurun_out/r02_bench2_$label.log 2>&1; echo "$label rc=$?"; grep '^{' gpurun_out/r02_bench2_$label.log | tail -1 | cut -c1-400; }
run default --steps 20 --warmup 5
run nosweepwait --steps 300 --warmup 5 --no-e2e --no-cfg3 --no-check --no-sweep-wait
run sweepwait --steps 300 --warmup 5 --no-e2e --no-cfg3 --no-check
run nccl --steps 300 --warmup 5 --no-e2e --no-cfg3 --no-check --no-p2p
run cfg4_p2p --workload cfg4 --steps 100 --warmup 5 --no-e2e
run cfg4_nccl --workload cfg4 --steps 100 --warmup 5 --no-e2e --no-p2p
( time timeout 600 python bench.py --workload cfg4 --steps 100 --warmup 5 --no-e2e ) > gpurun_out/r02_bench1_cfg4.log 2>&1; grep '^{' gpurun_out/r02_bench1_cfg4.log | tail -1 | cut -c1-300
