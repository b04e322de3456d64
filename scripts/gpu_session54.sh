#!/bin/bash
TAG=${1:-s54}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_ivector.py tests/test_gpu_fullsize.py tests/test_gpu_kaldi_exact.py tests/test_gpu_edges.py -m gpu -q 2>&1 | tail -6 ) > gpurun_out/${TAG}_tests.log
( timeout 300 python bench.py --config C3 --steps 50 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c3.log
( timeout 600 python bench.py --config C5 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c5.log
echo done
