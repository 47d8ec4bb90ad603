#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1
( time timeout 600 python bench.py ) > gpurun_out/bench_default.log 2>&1
( time timeout 300 python bench.py --impl reference --steps 10 --warmup 3 ) > gpurun_out/bench_ref.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -4 gpurun_out/bench_default.log | cut -c1-300
# optional extras (uncomment as needed):
# timeout 300 python bench.py --workload cfg3-256 --steps 200 --no-cpu-baseline --no-e2e
# timeout 300 python bench.py --workload cfg1 --steps 2000 --no-cpu-baseline --no-e2e
# timeout 600 python bench.py --workload cfg4 --steps 100
# ncu --set full --clock-control none --import-source on -k regex:sweepKernel -s 4 -c 2 -f -o gpurun_out/prof_sweep python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e
# the compiled host above the C ABI (NOT YET RUN ON A GPU): state file in, 40 steps, dump vs oracle
( timeout 200 python tests/host_driver_parity.py ) > gpurun_out/host_driver.log 2>&1; tail -6 gpurun_out/host_driver.log
