#!/bin/bash
# round 2, call 4 (1 GPU): GPU suite, the driver's N = 1 bench line, reference arm, cfg4 with the tiled
# interpolation, ncu launch lists and --set full captures (traffic of the sweep, the tile kernel)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r02_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
tail -14 gpurun_out/r02_pytest_gpu.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02_bench1_default.log 2>&1; echo "bench default rc=$?"
grep '^{' gpurun_out/r02_bench1_default.log | tail -1 | cut -c1-300
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r02_bench1_ref.log 2>&1; echo "ref rc=$?"
grep '^{' gpurun_out/r02_bench1_ref.log | tail -1 | cut -c1-300
( time timeout 600 python bench.py --steps 500 --warmup 5 --no-cfg3 --no-cpu-baseline --no-e2e ) > gpurun_out/r02_bench1_long.log 2>&1
grep '^{' gpurun_out/r02_bench1_long.log | tail -1 | cut -c1-300
( time timeout 600 python bench.py --workload cfg4 --steps 100 --warmup 5 --no-e2e ) > gpurun_out/r02_bench1_cfg4.log 2>&1
grep '^{' gpurun_out/r02_bench1_cfg4.log | tail -1 | cut -c1-300
( time timeout 600 python bench.py --workload cfg1 --steps 2000 --warmup 5 --no-e2e --no-cpu-baseline ) > gpurun_out/r02_bench1_cfg1.log 2>&1
grep '^{' gpurun_out/r02_bench1_cfg1.log | tail -1 | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_cfg2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-cfg3 > gpurun_out/r02_ncu_list_cfg2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/r02_launches_cfg4.csv python bench.py --workload cfg4 --steps 5 --warmup 3 --no-e2e > gpurun_out/r02_ncu_list_cfg4.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sweepKernel -s 4 -c 1 -f -o gpurun_out/r02_prof_sweep_trt19 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-cfg3 > gpurun_out/r02_ncu_sweep19.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sweepKernel -s 4 -c 1 -f -o gpurun_out/r02_prof_sweep_mrt27 python bench.py --workload cfg3-256 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_sweep27.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:intpTileKernel -s 2 -c 1 -f -o gpurun_out/r02_prof_intp_tile python bench.py --workload cfg4 --steps 3 --warmup 3 --no-e2e > gpurun_out/r02_ncu_intp.log 2>&1
ls -la gpurun_out/*.ncu-rep
