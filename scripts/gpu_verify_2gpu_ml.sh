#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_multirank.py -m gpu -x -q -k multilevel ) > gpurun_out/pytest_multi_ml.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_multi_ml.log
tail -30 gpurun_out/pytest_multi_ml.log
