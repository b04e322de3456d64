#!/bin/bash
# 8-GPU correctness: S-sharded attack (peer-memory exchange and ncclAllReduce) vs the single-GPU trajectory, 40 iterations.
TAG=${1:-s50}
N=${2:-8}
mkdir -p gpurun_out
( MG_ITERS=40 MG_S=48 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/multi_gpu_check.py 2>&1 | grep -E "^\[|^rank|p2p|MULTI|Error|error|assert" | tail -40 ) > gpurun_out/${TAG}_mg_check_${N}gpu.log
echo done
