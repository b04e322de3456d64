#!/bin/bash
# 8-GPU session: BASELINE.json configs[3] (C4) and configs[4] (C5) plus the metric's config (C2) at 8 ranks.
TAG=${1:-s9}
N=${2:-8}
mkdir -p gpurun_out
run() { # name port args...
  local name=$1 port=$2; shift 2
  ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@" 2>&1 | tail -3 ) > gpurun_out/${TAG}_bench_${name}_${N}gpu.log
}
run c4 29521 --config C4 --steps 30 --warmup 5
run c5 29522 --config C5 --steps 5 --warmup 3
run c2 29523 --config C2 --steps 50 --warmup 10
echo done
