#!/bin/bash
# round 2, call 1 (1 GPU): full GPU suite incl. the 128^3 / 256^3 oracle comparisons, the compiled host driver
mkdir -p gpurun_out
nproc > gpurun_out/r02_host.txt; free -g >> gpurun_out/r02_host.txt; nvidia-smi -L >> gpurun_out/r02_host.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r02_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
( timeout 300 python tests/host_driver_parity.py ) > gpurun_out/r02_host_driver.log 2>&1
echo "host_driver rc=$?" >> gpurun_out/r02_host_driver.log
tail -25 gpurun_out/r02_pytest_gpu.log; tail -8 gpurun_out/r02_host_driver.log
