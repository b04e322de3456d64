#!/bin/bash
TAG=${1:-s56}
LIB=${2:-q8}
mkdir -p gpurun_out
export FB_LIB_PATH=$PWD/fakebob_b200/libfb_$LIB.so
( timeout 600 python -m pytest tests/test_gpu_ivector.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -4 ) > gpurun_out/${TAG}_tests.log
( timeout 300 python bench.py --config C3 --steps 50 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c3.log
echo done
