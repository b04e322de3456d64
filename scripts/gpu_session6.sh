#!/bin/bash
# 2-GPU session: peer-memory exchange check + C2/C4 bench at 2 GPUs; plus single-GPU bench for the same build
TAG=${1:-s6}
mkdir -p gpurun_out
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py 2>&1 | tail -40 ) > gpurun_out/${TAG}_mg_check.log
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 10 2>&1 | tail -3 ) > gpurun_out/${TAG}_bench_c2_2gpu.log
( FB_NO_P2P=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 10 2>&1 | tail -3 ) > gpurun_out/${TAG}_bench_c2_2gpu_nccl.log
( timeout 300 python bench.py --steps 50 --warmup 10 --no-extra 2>&1 | tail -2 ) > gpurun_out/${TAG}_bench_c2.log
( timeout 300 python -m pytest tests/test_gpu_gmm.py tests/test_gpu_fullsize.py tests/test_gpu_nes.py -m gpu -q 2>&1 | tail -15 ) > gpurun_out/${TAG}_tests.log
echo done
