#!/bin/bash
# round 2, 2 GPUs: overlapped exchange, single launch with the halo CTAs appended at the end; tile kernel with unrolled staging
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_multirank.py -m gpu -q -k "overlap or timeout or sweepwait or (p2p and not multilevel)" ) > gpurun_out/r02_pytest_multi12.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_multi12.log; tail -6 gpurun_out/r02_pytest_multi12.log
( time timeout 600 python -m pytest tests/test_multilevel.py tests/test_coupled_multilevel.py -m gpu -x -q ) > gpurun_out/r02_pytest_gpu12.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu12.log; tail -3 gpurun_out/r02_pytest_gpu12.log
two() { label=$1; shift; ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 2 "$@" ) > gpurun_out/r02_h2_$label.log 2>&1; echo "$label rc=$?"; grep '^{' gpurun_out/r02_h2_$label.log | tail -1 | cut -c1-200; }
two overlap --steps 300 --warmup 5 --no-e2e --no-cfg3 --no-check
two nooverlap --steps 300 --warmup 5 --no-e2e --no-cfg3 --no-check --no-overlap
two overlap_b --steps 300 --warmup 5 --no-e2e --no-cfg3 --no-check
two nooverlap_b --steps 300 --warmup 5 --no-e2e --no-cfg3 --no-check --no-overlap
two cfg3_overlap --workload cfg3 --steps 60 --warmup 5 --no-e2e
two cfg3_nooverlap --workload cfg3 --steps 60 --warmup 5 --no-e2e --no-overlap
( time timeout 600 python bench.py --workload cfg4 --steps 100 --warmup 5 --no-e2e ) > gpurun_out/r02_h1_cfg4.log 2>&1; grep '^{' gpurun_out/r02_h1_cfg4.log | tail -1 | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/r02_launches_cfg4.csv python bench.py --workload cfg4 --steps 5 --warmup 3 --no-e2e > gpurun_out/r02_ncu_list_cfg4.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:intpTileKernel -s 2 -c 1 -f -o gpurun_out/r02_prof_intp_tile python bench.py --workload cfg4 --steps 3 --warmup 3 --no-e2e > gpurun_out/r02_ncu_intp.log 2>&1
