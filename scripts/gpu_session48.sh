#!/bin/bash
TAG=${1:-s48}
KERNEL=${2:-mfcc_kernel}
mkdir -p gpurun_out
export FB_NO_GRAPH=1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name regex:$KERNEL --launch-skip 1 --launch-count 1 -o gpurun_out/${TAG}_k python scripts/profile_iter.py 2 C2 > gpurun_out/${TAG}_ncu.log 2>&1
echo done
