#!/bin/bash
# usage: exp_multi.sh <ngpus> <steps> <label> [bench args...]   (env vars are inherited)
n=$1; steps=$2; label=$3; shift 3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $n --steps $steps --warmup 5 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); t=d['timers_ms_per_step']
print('$label: MLUPS %.0f ms/step %.4f frac %.3f compute %.4f bc %.4f comm %.4f clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], t['compute'], t['bc'], t['comm'], d['clocks']['sm_mhz']))"
