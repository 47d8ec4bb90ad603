#!/bin/bash
# round 2 experiment (profiles/r02_small_mesh_cache_policy.md; the sweep_keep.cu TU it needs is quoted there, not shipped):
# L2-friendly build of the sweep for small meshes -- parity forced on for
# every level, then cfg1 (64^3) and 32^3 / 128^3 with the default kernel and with the keep kernel
mkdir -p gpurun_out
( MUSB200_KEEP_L2_MB=1000000 timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_multilevel.py tests/test_gpu_golden.py -m gpu -x -q ) > gpurun_out/r02_keep_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_keep_pytest.log; tail -3 gpurun_out/r02_keep_pytest.log
for lvl in 5 6 7; do for mb in 0 1000000; do
  echo "level $lvl keep_mb $mb" >> gpurun_out/r02_keep_bench.log
  MUSB200_KEEP_L2_MB=$mb timeout 100 python bench.py --workload cfg1 --level $lvl --steps 4000 --warmup 200 --no-e2e --no-cpu-baseline --no-check >> gpurun_out/r02_keep_bench.log 2>&1
done; done
grep -o 'level [0-9] keep_mb [0-9]*\|"value": [0-9.]*, "unit": "MLUPS", "n_gpus"\|"ms_per_step": [0-9.]*' gpurun_out/r02_keep_bench.log
