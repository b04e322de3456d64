#!/bin/bash
TAG=${1:-s42}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_ivector.py -m gpu -q -x -k "reproducible" 2>&1 | tail -40 ) > gpurun_out/${TAG}_tests.log
echo done
