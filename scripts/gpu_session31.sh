#!/bin/bash
TAG=${1:-s31}
mkdir -p gpurun_out
export FB_LIB_PATH=$PWD/fakebob_b200/libfb_solvestats.so
( timeout 300 python bench.py --config C3 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep "solve " | tail -6 ) > gpurun_out/${TAG}_solvestats.log
export FB_LIB_PATH=$PWD/fakebob_b200/libfb_pq.so
( timeout 600 python -m pytest tests/test_gpu_ivector.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -5 ) > gpurun_out/${TAG}_tests.log
( timeout 300 python bench.py --config C3 --steps 50 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c3.log
echo done
