#!/bin/bash
TAG=${1:-s46}
LIB=${2:-fe}
mkdir -p gpurun_out
( timeout 400 python bench.py --steps 50 --warmup 10 --no-extra --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c2_base.log
export FB_LIB_PATH=$PWD/fakebob_b200/libfb_$LIB.so
( timeout 900 python -m pytest tests/test_gpu_gmm.py tests/test_gpu_fullsize.py tests/test_gpu_edges.py -m gpu -q 2>&1 | tail -6 ) > gpurun_out/${TAG}_tests.log
( timeout 400 python bench.py --steps 50 --warmup 10 --no-extra --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c2_$LIB.log
echo done
