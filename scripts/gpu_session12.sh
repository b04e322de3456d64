#!/bin/bash
TAG=${1:-s12}
mkdir -p gpurun_out
export FB_LIB_PATH=$PWD/fakebob_b200/libfb_post2.so
( timeout 900 python -m pytest tests/test_gpu_ivector.py tests/test_gpu_fullsize.py tests/test_gpu_kaldi_exact.py -m gpu -q 2>&1 | tail -15 ) > gpurun_out/${TAG}_tests.log
( timeout 400 python bench.py --config C3 --steps 50 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c3_post2.log
unset FB_LIB_PATH
( timeout 400 python bench.py --config C3 --steps 50 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c3_base.log
echo done
