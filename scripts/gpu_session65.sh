#!/bin/bash
TAG=${1:-s65}
mkdir -p gpurun_out
( timeout 150 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/${TAG}_smoke.log
( timeout 200 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c2.log
echo done
