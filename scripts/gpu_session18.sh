#!/bin/bash
TAG=${1:-s18}
LIB=${2:-pq}
mkdir -p gpurun_out
export FB_LIB_PATH=$PWD/fakebob_b200/libfb_$LIB.so
export FB_NO_GRAPH=1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on --launch-skip 18 --launch-count 18 -o gpurun_out/${TAG}_c3 python scripts/profile_iter.py 2 C3 > gpurun_out/${TAG}_ncu_c3.log 2>&1
echo done
