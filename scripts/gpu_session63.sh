#!/bin/bash
TAG=${1:-s63}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q -rA 2>&1 | tail -100 ) > gpurun_out/${TAG}_tests.log
( timeout 400 python bench.py --config C3 --steps 50 --warmup 10 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c3.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/${TAG}_smoke.log
echo done
