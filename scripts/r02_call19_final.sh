#!/bin/bash
# round 2, last 1-GPU pass after the sweep's cache hints changed (plain ld.global.nc / st.global instead of
# .cs): GPU suite, traffic captures of the final kernel, the headline line
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_baseline_sizes.py ) > gpurun_out/r02_pytest_gpu_final2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_final2.log; tail -4 gpurun_out/r02_pytest_gpu_final2.log
timeout 200 ncu --set full --clock-control none -k regex:sweepKernel -s 4 -c 1 -f -o gpurun_out/r02b_prof_sweep_trt19 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-cfg3 > gpurun_out/r02b_ncu_sweep19.log 2>&1; echo "ncu19 rc=$?"
timeout 200 ncu --set full --clock-control none -k regex:sweepKernel -s 4 -c 1 -f -o gpurun_out/r02b_prof_sweep_mrt27 python bench.py --workload cfg3-256 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02b_ncu_sweep27.log 2>&1; echo "ncu27 rc=$?"
( timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-cfg3 ) > gpurun_out/r02b_bench20.log 2>&1; grep '^{' gpurun_out/r02b_bench20.log | tail -1 | cut -c1-420
( timeout 200 python bench.py --steps 500 --warmup 5 --no-cpu-baseline --no-e2e --no-cfg3 ) > gpurun_out/r02b_bench500.log 2>&1; grep '^{' gpurun_out/r02b_bench500.log | tail -1 | cut -c1-420
