#!/bin/bash
TAG=${1:-s41}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_ivector.py tests/test_gpu_fullsize.py tests/test_gpu_kaldi_exact.py -m gpu -q 2>&1 | tail -8 ) > gpurun_out/${TAG}_tests.log
( timeout 300 python bench.py --config C3 --steps 50 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_c3.log
( timeout 900 compute-sanitizer --tool racecheck --print-limit 3 python -m pytest tests/test_gpu_ivector.py -m gpu -q -x -k "reproducible" 2>&1 | grep -E "RACECHECK SUMMARY|passed|failed|Error: Race" | sort | uniq -c | head -20 ) > gpurun_out/${TAG}_racecheck_iv.log
echo done
