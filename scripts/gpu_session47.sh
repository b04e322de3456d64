#!/bin/bash
TAG=${1:-s47}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q -rA 2>&1 | tail -100 ) > gpurun_out/${TAG}_tests.log
( timeout 400 python bench.py --steps 50 --warmup 10 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c2.log
( timeout 400 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c2_driver_args.log
( timeout 400 python bench.py --config C3 --steps 50 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c3.log
( timeout 400 python bench.py --config C4 --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c4.log
echo done
