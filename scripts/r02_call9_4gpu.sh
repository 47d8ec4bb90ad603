#!/bin/bash
# round 2, 8 GPUs: the driver's own invocation at N = 8 (weak-scaled cavity + multirank check + cfg3 512^3
# strong-scaled), 8-octant and multi-level parity against the oracle, exchange variants, cfg4 / cfg5
mkdir -p gpurun_out
N=4
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 "$@"; }
run() { label=$1; shift; ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $N "$@" ) > gpurun_out/r02_e${N}_$label.log 2>&1; echo "$label rc=$?"; grep '^{' gpurun_out/r02_e${N}_$label.log | tail -1 | cut -c1-220; }
run default --steps 20 --warmup 5
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29651 tests/parity_multi.py --mode gpu --layout d3q19 --relaxation trt --kind cavity --level 5 --steps 40 --octants 4 --p2p ) > gpurun_out/r02_parity4_cavity_p2p.log 2>&1; grep -c "ndiff=0" gpurun_out/r02_parity4_cavity_p2p.log
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29652 tests/parity_multi.py --mode gpu --layout d3q27 --relaxation mrt --kind periodic --level 5 --steps 40 --p2p ) > gpurun_out/r02_parity4_mrt27_p2p.log 2>&1; grep -c "ndiff=0" gpurun_out/r02_parity4_mrt27_p2p.log
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29653 tests/parity_multi.py --mode gpu --layout d3q19 --relaxation bgk --kind channel --level 5 --steps 40 ) > gpurun_out/r02_parity4_channel_nccl.log 2>&1; grep -c "ndiff=0" gpurun_out/r02_parity4_channel_nccl.log
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29654 tests/parity_multi.py --mode gpu-ml --layout d3q19 --relaxation bgk --levels 3 --method linear --steps 10 --p2p ) > gpurun_out/r02_parity4_ml3_p2p.log 2>&1; grep "multilevel" gpurun_out/r02_parity4_ml3_p2p.log | grep -c "ndiff=0"
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 tests/parity_multi.py --mode gpu-ml --layout d3q19 --relaxation bgk --levels 2 --method linear --steps 10 --restart --p2p ) > gpurun_out/r02_parity4_ml2_restart.log 2>&1; grep "multilevel" gpurun_out/r02_parity4_ml2_restart.log | grep -c "ndiff=0"
run nosweepwait --steps 300 --warmup 5 --no-e2e --no-cfg3 --no-check --no-sweep-wait
run sweepwait --steps 300 --warmup 5 --no-e2e --no-cfg3 --no-check
run cfg4 --workload cfg4 --steps 60 --warmup 5 --no-e2e
run cfg5 --workload cfg5 --steps 60 --warmup 5 --no-e2e
