#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:intpKernel -s 2 -c 1 -f -o gpurun_out/prof_intp python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_intp.log 2>&1
( time timeout 600 python -m pytest tests/test_multilevel.py tests/test_restart.py -m gpu -x -q ) > gpurun_out/pytest_ml.log 2>&1
timeout 600 python bench.py --workload cfg4 --steps 100 --no-e2e > gpurun_out/bench_cfg4.log 2>&1
tail -3 gpurun_out/pytest_ml.log; tail -1 gpurun_out/bench_cfg4.log | cut -c1-200
