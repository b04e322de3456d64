#!/bin/bash
# Profiles (ncu launch lists + full captures) and compute-sanitizer runs.  Outputs under gpurun_out/<tag>_*.
TAG=${1:-s7}
mkdir -p gpurun_out
export FB_NO_GRAPH=1
# launch lists of one NES iteration (second of two iterations inside the profiler range), C2 and C3
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_c2.csv python scripts/profile_iter.py 2 C2 > gpurun_out/${TAG}_prof_c2.log 2>&1
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_c3.csv python scripts/profile_iter.py 2 C3 > gpurun_out/${TAG}_prof_c3.log 2>&1
# full captures: every kernel of the second C2 iteration (8 kernels), and of the second C3 iteration
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on --launch-skip 8 --launch-count 8 -o gpurun_out/${TAG}_c2 python scripts/profile_iter.py 2 C2 > gpurun_out/${TAG}_ncu_c2.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on --launch-skip 18 --launch-count 18 -o gpurun_out/${TAG}_c3 python scripts/profile_iter.py 2 C3 > gpurun_out/${TAG}_ncu_c3.log 2>&1
unset FB_NO_GRAPH
# compute-sanitizer: the smoke run (GMM + i-vector path incl. tcgen05 / mbarrier / cluster kernels)
for tool in memcheck racecheck synccheck; do
  ( timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -40 ) > gpurun_out/${TAG}_sanitizer_${tool}.log
done
( timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_ivector.py tests/test_gpu_edges.py -m gpu -q -x 2>&1 | tail -30 ) > gpurun_out/${TAG}_sanitizer_memcheck_tests.log
echo done
