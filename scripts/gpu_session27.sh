#!/bin/bash
# Round-2 refresh after the i-vector kernel work: full GPU suite, benches of all configs on one GPU, reference arms, profiles.
TAG=${1:-s27}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q -rA 2>&1 | tail -100 ) > gpurun_out/${TAG}_tests.log
( timeout 400 python bench.py --steps 50 --warmup 10 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c2.log
( timeout 400 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c2_driver_args.log
( timeout 400 python bench.py --config C3 --steps 50 --warmup 10 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c3.log
( timeout 400 python bench.py --config C4 --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c4.log
( timeout 600 python bench.py --config C5 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c5.log
export FB_NO_GRAPH=1
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_c3.csv python scripts/profile_iter.py 2 C3 > gpurun_out/${TAG}_prof_c3.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on --launch-skip 18 --launch-count 18 -o gpurun_out/${TAG}_c3 python scripts/profile_iter.py 2 C3 > gpurun_out/${TAG}_ncu_c3.log 2>&1
echo done
