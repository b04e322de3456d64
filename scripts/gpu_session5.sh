#!/bin/bash
TAG=${1:-s5}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -rA 2>&1 ) > gpurun_out/${TAG}_tests.log
( timeout 400 python bench.py --steps 50 --warmup 10 2>&1 | tail -2 ) > gpurun_out/${TAG}_bench_c2.log
( timeout 400 python bench.py --config C4 --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -2 ) > gpurun_out/${TAG}_bench_c4.log
( timeout 600 python bench.py --config C5 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 ) > gpurun_out/${TAG}_bench_c5.log
echo done
