#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 300 > gpurun_out/bench_n1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sweepKernel -s 4 -c 2 -f -o gpurun_out/prof_sweep_trt19 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_trt19.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/bench_n1.log | cut -c1-400
