#!/bin/bash
TAG=${1:-s11}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q -rA 2>&1 | tail -100 ) > gpurun_out/${TAG}_tests.log
( timeout 400 python bench.py --steps 50 --warmup 10 2>&1 | tail -2 ) > gpurun_out/${TAG}_bench_c2.log
( timeout 400 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline 2>&1 | tail -2 ) > gpurun_out/${TAG}_bench_c2_short.log
( FB_GMM_GENERIC_SLOTS=1 timeout 400 python bench.py --steps 50 --warmup 10 --no-extra --no-cpu-baseline 2>&1 | tail -2 ) > gpurun_out/${TAG}_bench_c2_generic_slots.log
( timeout 400 python bench.py --config C3 --steps 50 --warmup 10 --no-cpu-baseline 2>&1 | tail -2 ) > gpurun_out/${TAG}_bench_c3.log
echo done
