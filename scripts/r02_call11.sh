#!/bin/bash
# round 2, 2 GPUs: overlapped exchange (push on a second stream + CTA-granular split sweep): parity, timing
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_multirank.py -m gpu -x -q -k "overlap or timeout or sweepwait or (p2p and not multilevel)" ) > gpurun_out/r02_pytest_multi11.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_multi11.log; tail -6 gpurun_out/r02_pytest_multi11.log
two() { label=$1; shift; ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 2 "$@" ) > gpurun_out/r02_g2_$label.log 2>&1; echo "$label rc=$?"; grep '^{' gpurun_out/r02_g2_$label.log | tail -1 | cut -c1-200; }
two overlap --steps 300 --warmup 5 --no-e2e --no-cfg3 --no-check
two nooverlap --steps 300 --warmup 5 --no-e2e --no-cfg3 --no-check --no-overlap
two overlap_b --steps 300 --warmup 5 --no-e2e --no-cfg3 --no-check
two nooverlap_b --steps 300 --warmup 5 --no-e2e --no-cfg3 --no-check --no-overlap
two cfg3_overlap --workload cfg3 --steps 60 --warmup 5 --no-e2e
two cfg3_nooverlap --workload cfg3 --steps 60 --warmup 5 --no-e2e --no-overlap
two default --steps 20 --warmup 5
bash scripts/r02_call10.sh
