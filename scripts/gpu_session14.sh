#!/bin/bash
TAG=${1:-s14}
LIB=${2:-post3}
mkdir -p gpurun_out
( timeout 120 scripts/fp64_probe 2>&1 ) > gpurun_out/${TAG}_fp64_probe.log
export FB_LIB_PATH=$PWD/fakebob_b200/libfb_$LIB.so
( timeout 600 python -m pytest tests/test_gpu_ivector.py tests/test_gpu_fullsize.py tests/test_gpu_kaldi_exact.py -m gpu -q -x 2>&1 | tail -25 ) > gpurun_out/${TAG}_tests.log
( timeout 300 python bench.py --config C3 --steps 50 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c3_$LIB.log
export FB_NO_GRAPH=1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on --launch-skip 18 --launch-count 18 -o gpurun_out/${TAG}_c3 python scripts/profile_iter.py 2 C3 > gpurun_out/${TAG}_ncu_c3.log 2>&1
echo done
