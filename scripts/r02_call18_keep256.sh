#!/bin/bash
# round 2 experiment, part 2 (profiles/r02_small_mesh_cache_policy.md): the cache policy at 256^3
mkdir -p gpurun_out
for wl in cfg2 cfg3-256; do for mb in 0 1000000; do
  echo "workload $wl keep_mb $mb" >> gpurun_out/r02_keep256.log
  MUSB200_KEEP_L2_MB=$mb timeout 150 python bench.py --workload $wl --steps 400 --warmup 20 --no-e2e --no-cpu-baseline --no-check --no-cfg3 >> gpurun_out/r02_keep256.log 2>&1
done; done
grep -o 'workload [a-z0-9-]* keep_mb [0-9]*\|"value": [0-9.]*, "unit": "MLUPS", "n_gpus"\|"ms_per_step": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r02_keep256.log
