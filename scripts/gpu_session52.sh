#!/bin/bash
TAG=${1:-s52}
mkdir -p gpurun_out
for LIB in base fe1 fe2; do
  if [ "$LIB" == "base" ]; then unset FB_LIB_PATH; else export FB_LIB_PATH=$PWD/fakebob_b200/libfb_$LIB.so; fi
  ( timeout 300 python bench.py --steps 50 --warmup 10 --no-extra --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c2_$LIB.log
done
export FB_LIB_PATH=$PWD/fakebob_b200/libfb_fe1.so
( timeout 600 python -m pytest tests/test_gpu_gmm.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -4 ) > gpurun_out/${TAG}_tests_fe1.log
echo done
