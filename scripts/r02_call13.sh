#!/bin/bash
# round 2, final 1-GPU pass: whole GPU suite, the driver's N = 1 invocation, traffic captures of the final kernel
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/r02_pytest_gpu_final.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_final.log; tail -12 gpurun_out/r02_pytest_gpu_final.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02_final1_default.log 2>&1; echo "bench default rc=$?"
grep '^{' gpurun_out/r02_final1_default.log | tail -1 | cut -c1-300
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/r02_final_smoke.log
( time timeout 600 python bench.py --workload cfg4 --steps 100 --warmup 5 --no-e2e ) > gpurun_out/r02_final1_cfg4.log 2>&1; grep '^{' gpurun_out/r02_final1_cfg4.log | tail -1 | cut -c1-200
( time timeout 600 python bench.py --workload cfg5 --steps 60 --warmup 5 --no-e2e ) > gpurun_out/r02_final1_cfg5.log 2>&1; grep '^{' gpurun_out/r02_final1_cfg5.log | tail -1 | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_cfg2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-cfg3 > gpurun_out/r02_ncu_list_cfg2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sweepKernel -s 4 -c 1 -f -o gpurun_out/r02_prof_sweep_trt19 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-cfg3 > gpurun_out/r02_ncu_sweep19.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sweepKernel -s 4 -c 1 -f -o gpurun_out/r02_prof_sweep_mrt27 python bench.py --workload cfg3-256 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_sweep27.log 2>&1
