#!/bin/bash
TAG=${1:-s24}
LIB=${2:-pq}
KERNEL=${3:-ivec_solve_kernel}
mkdir -p gpurun_out
export FB_LIB_PATH=$PWD/fakebob_b200/libfb_$LIB.so
export FB_NO_GRAPH=1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name regex:$KERNEL --launch-skip 1 --launch-count 1 -o gpurun_out/${TAG}_k python scripts/profile_iter.py 2 C3 > gpurun_out/${TAG}_ncu.log 2>&1
echo done
