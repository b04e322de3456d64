#!/bin/bash
TAG=${1:-s8}
mkdir -p gpurun_out
bash scripts/gmm_ab.sh ${TAG}
( timeout 900 python -m pytest tests -m gpu -q -rA 2>&1 | tail -80 ) > gpurun_out/${TAG}_tests.log
( timeout 300 python bench.py --steps 50 --warmup 10 --no-extra 2>&1 | tail -2 ) > gpurun_out/${TAG}_bench.log
echo done
