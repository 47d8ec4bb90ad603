#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload cfg4 --steps 100 > gpurun_out/bench_cfg4.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 70 --csv --log-file gpurun_out/launches_cfg4.csv python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list4.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/bench_cfg4.log | cut -c1-300
