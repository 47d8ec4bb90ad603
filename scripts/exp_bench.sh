#!/bin/bash
# usage: exp_bench.sh <workload> <steps> variant...   ('a' = the in-tree build)
wl=$1; steps=$2; shift 2
for v in "$@"; do
  if [ $v = a ]; then unset MUSB200_LIB; else export MUSB200_LIB=$PWD/exp/lib_$v.so; fi
  python bench.py --workload $wl --steps $steps --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$wl variant $v: MLUPS %.0f kernel_ms %.4f frac %.3f clocks %s %s' % (d['value'], r['kernel_ms'], r['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"
done
