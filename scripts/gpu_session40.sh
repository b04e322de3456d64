#!/bin/bash
# Round-2 final refresh on one GPU: full GPU suite, bench lines of every config, launch lists, full ncu captures, sanitizer.
TAG=${1:-s40}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q -rA 2>&1 | tail -100 ) > gpurun_out/${TAG}_tests.log
( timeout 400 python bench.py --steps 50 --warmup 10 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c2.log
( timeout 400 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c2_driver_args.log
( timeout 400 python bench.py --config C3 --steps 50 --warmup 10 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c3.log
( timeout 400 python bench.py --config C4 --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c4.log
( timeout 600 python bench.py --config C5 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c5.log
export FB_NO_GRAPH=1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_c2.csv python scripts/profile_iter.py 2 C2 > gpurun_out/${TAG}_prof_c2.log 2>&1
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_c3.csv python scripts/profile_iter.py 2 C3 > gpurun_out/${TAG}_prof_c3.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on --launch-skip 10 --launch-count 9 -o gpurun_out/${TAG}_c2 python scripts/profile_iter.py 2 C2 > gpurun_out/${TAG}_ncu_c2.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on --launch-skip 16 --launch-count 15 -o gpurun_out/${TAG}_c3 python scripts/profile_iter.py 2 C3 > gpurun_out/${TAG}_ncu_c3.log 2>&1
unset FB_NO_GRAPH
for tool in memcheck racecheck synccheck; do
  ( timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -40 ) > gpurun_out/${TAG}_sanitizer_${tool}.log
done
( timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_ivector.py tests/test_gpu_edges.py -m gpu -q -x 2>&1 | tail -30 ) > gpurun_out/${TAG}_sanitizer_memcheck_tests.log
echo done
