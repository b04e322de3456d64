#!/bin/bash
# Same-box A/B of GMM kernel variants (built beforehand with scripts/build_variant.sh); output gpurun_out/<tag>_gmm_ab.log
TAG=${1:-ab}
export FB_TREE_ROOT=/tmp/fb_tree_ab
mkdir -p $FB_TREE_ROOT gpurun_out
{
python scripts/gmm_time.py
for lib in "" $(ls fakebob_b200/libfb_*.so 2>/dev/null); do
  for t in 2; do
    echo "== lib=${lib:-default} terms=$t"
    if [[ "$lib" == *stats* ]]; then
      FB_LIB_PATH=$PWD/$lib FAKEBOB_GMM_DELTA_TERMS=$t python scripts/gmm_stats.py
    else
      FB_LIB_PATH=${lib:+$PWD/$lib} FAKEBOB_GMM_DELTA_TERMS=$t python scripts/gmm_time.py
    fi
  done
done
} > gpurun_out/${TAG}_gmm_ab.log 2>&1
