#!/bin/bash
# round 2, 2 GPUs: exchange timeout tests, compute-sanitizer on the exchange kernels and the new tile kernel
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_multirank.py -m gpu -x -q -k "timeout" ) > gpurun_out/r02_pytest_timeout.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_timeout.log; tail -5 gpurun_out/r02_pytest_timeout.log
( timeout 900 compute-sanitizer --tool memcheck --target-processes all --log-file gpurun_out/r02_memcheck_p2p.%p.log python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29661 tests/parity_multi.py --mode gpu --layout d3q19 --relaxation trt --kind cavity --level 4 --steps 12 --octants 2 --p2p ) > gpurun_out/r02_memcheck_p2p.out 2>&1
echo "memcheck p2p rc=$?"; grep -h "ERROR SUMMARY\|ndiff" gpurun_out/r02_memcheck_p2p.*log gpurun_out/r02_memcheck_p2p.out | tail -6
( timeout 900 compute-sanitizer --tool memcheck --target-processes all --log-file gpurun_out/r02_memcheck_ml.%p.log python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29662 tests/parity_multi.py --mode gpu-ml --layout d3q19 --relaxation bgk --levels 2 --method linear --steps 4 --p2p ) > gpurun_out/r02_memcheck_ml.out 2>&1
echo "memcheck ml rc=$?"; grep -h "ERROR SUMMARY\|multilevel" gpurun_out/r02_memcheck_ml.*log gpurun_out/r02_memcheck_ml.out | tail -6
( timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r02_racecheck_tile.log python -m pytest tests/test_multilevel.py tests/test_coupled_multilevel.py -m gpu -x -q -k "(2lvl-linear-bgk19 or 2lvl-quad or 2lvl-wavg) and not target-major" ) > gpurun_out/r02_racecheck_tile.out 2>&1
echo "racecheck rc=$?"; grep -h "RACECHECK SUMMARY\|passed\|failed" gpurun_out/r02_racecheck_tile.log gpurun_out/r02_racecheck_tile.out | tail -4
( timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_memcheck_1gpu.log python -m pytest tests/test_gpu_cube.py tests/test_restart.py tests/test_coupled_multilevel.py -m gpu -x -q -k "not 6-" ) > gpurun_out/r02_memcheck_1gpu.out 2>&1
echo "memcheck 1gpu rc=$?"; grep -h "ERROR SUMMARY\|passed\|failed" gpurun_out/r02_memcheck_1gpu.log gpurun_out/r02_memcheck_1gpu.out | tail -4
( time timeout 600 python bench.py --workload cfg4 --steps 100 --warmup 5 --no-e2e ) > gpurun_out/r02_f1_cfg4.log 2>&1; grep '^{' gpurun_out/r02_f1_cfg4.log | tail -1 | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/r02_launches_cfg4.csv python bench.py --workload cfg4 --steps 5 --warmup 3 --no-e2e > gpurun_out/r02_ncu_list_cfg4.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:intpTileKernel -s 2 -c 1 -f -o gpurun_out/r02_prof_intp_tile python bench.py --workload cfg4 --steps 3 --warmup 3 --no-e2e > gpurun_out/r02_ncu_intp.log 2>&1
