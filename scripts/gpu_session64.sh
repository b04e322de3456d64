#!/bin/bash
# 8-GPU refresh of the two configurations whose kernels changed late in round 2: C5 (i-vector) and C2 (MFCC).
TAG=${1:-s64}
N=${2:-8}
mkdir -p gpurun_out
run() { local name=$1 port=$2; shift 2
  ( timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@" 2>&1 | tail -3 ) > gpurun_out/${TAG}_bench_${name}_${N}gpu.log
}
run c5 29542 --config C5 --steps 5 --warmup 3 --no-cpu-baseline
run c2 29543 --config C2 --steps 50 --warmup 10 --no-cpu-baseline --no-extra
echo done
