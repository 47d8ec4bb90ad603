#!/bin/bash
# round 2, last call (1 GPU): the GPU suite of the final tree, fast subset first
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_baseline_sizes.py ) > gpurun_out/r02_pytest_gpu_last.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_last.log; tail -6 gpurun_out/r02_pytest_gpu_last.log
