#!/bin/bash
# One gpurun call: GPU tests, bench lines, contraction A/B, ncu launch list.  Outputs under gpurun_out/<tag>_*.
TAG=${1:-s}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.log 2>&1
( timeout 1200 python -m pytest tests -m gpu -q -rA --durations=15 2>&1 | tail -150 ) > gpurun_out/${TAG}_tests.log
( timeout 400 python bench.py --steps 50 --warmup 10 2>&1 | tail -5 ) > gpurun_out/${TAG}_bench.log
for t in 1 2 3; do
  ( FAKEBOB_GMM_DELTA_TERMS=$t timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra 2>&1 | tail -2 ) > gpurun_out/${TAG}_bench_terms$t.log
done
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_ncu_stdout.log 2>&1 )
echo done
