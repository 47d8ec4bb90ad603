#!/bin/bash
# round 2, call 5 (2 GPUs): new GPU tests (coupled multi-level, tiled intp v2), cfg4 / cfg5 on 1 and 2 GPUs,
# cfg2 weak-scaled exchange variants with the faster push kernel
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_coupled_multilevel.py tests/test_multilevel.py tests/test_restart.py tests/test_sources.py -m gpu -x -q ) > gpurun_out/r02_pytest_gpu5.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu5.log; tail -6 gpurun_out/r02_pytest_gpu5.log
( time timeout 900 python -m pytest tests/test_multirank.py -m gpu -x -q -k "p2p or restart" ) > gpurun_out/r02_pytest_multi5.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_multi5.log; tail -4 gpurun_out/r02_pytest_multi5.log
one() { label=$1; shift; ( time timeout 600 python bench.py "$@" ) > gpurun_out/r02_b1_$label.log 2>&1; echo "$label rc=$?"; grep '^{' gpurun_out/r02_b1_$label.log | tail -1 | cut -c1-250; }
two() { label=$1; shift; ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 2 "$@" ) > gpurun_out/r02_b2_$label.log 2>&1; echo "$label rc=$?"; grep '^{' gpurun_out/r02_b2_$label.log | tail -1 | cut -c1-250; }
one cfg4 --workload cfg4 --steps 100 --warmup 5 --no-e2e
one cfg5 --workload cfg5 --steps 60 --warmup 5 --no-e2e
two cfg4 --workload cfg4 --steps 100 --warmup 5 --no-e2e
two cfg5 --workload cfg5 --steps 60 --warmup 5 --no-e2e
two sweepwait --steps 300 --warmup 5 --no-e2e --no-cfg3 --no-check
two nosweepwait --steps 300 --warmup 5 --no-e2e --no-cfg3 --no-check --no-sweep-wait
two sweepwait_b --steps 300 --warmup 5 --no-e2e --no-cfg3 --no-check
two nosweepwait_b --steps 300 --warmup 5 --no-e2e --no-cfg3 --no-check --no-sweep-wait
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/r02_launches_cfg4.csv python bench.py --workload cfg4 --steps 5 --warmup 3 --no-e2e > gpurun_out/r02_ncu_list_cfg4.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:intpTileKernel -s 2 -c 1 -f -o gpurun_out/r02_prof_intp_tile python bench.py --workload cfg4 --steps 3 --warmup 3 --no-e2e > gpurun_out/r02_ncu_intp.log 2>&1
