#!/bin/bash
TAG=${1:-s2}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_fullsize.py::test_c2_nes_trajectory_matches_fixture tests/test_gpu_gmm.py tests/test_gpu_ivector.py tests/test_gpu_nes.py tests/test_gpu_edges.py -q -rA 2>&1 ) > gpurun_out/${TAG}_tests.log
bash scripts/gmm_ab.sh ${TAG}
( timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-extra 2>&1 | tail -2 ) > gpurun_out/${TAG}_bench_pdl.log
( FB_NO_PDL=1 timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-extra 2>&1 | tail -2 ) > gpurun_out/${TAG}_bench_nopdl.log
echo done
