#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 400 python bench.py --steps 300 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1
timeout 300 python bench.py --workload cfg1 --steps 2000 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg1.log 2>&1
timeout 600 python bench.py --workload cfg4 --steps 100 > gpurun_out/bench_cfg4.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 70 --csv --log-file gpurun_out/launches_cfg4.csv python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list4.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; for f in n1 cfg1 cfg4; do tail -1 gpurun_out/bench_$f.log | cut -c1-200; done
