#!/bin/bash
# round 2, call 7 (1 GPU): tile kernel v4
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_multilevel.py tests/test_coupled_multilevel.py tests/test_restart.py tests/test_sources.py -m gpu -x -q ) > gpurun_out/r02_pytest_gpu7.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu7.log; tail -4 gpurun_out/r02_pytest_gpu7.log
( time timeout 600 python bench.py --workload cfg4 --steps 100 --warmup 5 --no-e2e ) > gpurun_out/r02_d1_cfg4.log 2>&1; grep '^{' gpurun_out/r02_d1_cfg4.log | tail -1 | cut -c1-200
( time timeout 600 python bench.py --workload cfg5 --steps 60 --warmup 5 --no-e2e ) > gpurun_out/r02_d1_cfg5.log 2>&1; grep '^{' gpurun_out/r02_d1_cfg5.log | tail -1 | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/r02_launches_cfg4.csv python bench.py --workload cfg4 --steps 5 --warmup 3 --no-e2e > gpurun_out/r02_ncu_list_cfg4.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:intpTileKernel -s 2 -c 1 -f -o gpurun_out/r02_prof_intp_tile python bench.py --workload cfg4 --steps 3 --warmup 3 --no-e2e > gpurun_out/r02_ncu_intp.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fromFinerFusedKernel -s 2 -c 1 -f -o gpurun_out/r02_prof_fromfiner python bench.py --workload cfg4 --steps 3 --warmup 3 --no-e2e > gpurun_out/r02_ncu_ff.log 2>&1
