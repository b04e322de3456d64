#!/bin/bash
TAG=${1:-s10}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -rA 2>&1 | tail -90 ) > gpurun_out/${TAG}_tests.log
( timeout 400 python bench.py --steps 20 --warmup 5 2>&1 | tail -2 ) > gpurun_out/${TAG}_bench_c2_driver_args.log
( timeout 400 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -2 ) > gpurun_out/${TAG}_bench_c2_reference.log
( timeout 400 python bench.py --config C3 --steps 50 --warmup 10 2>&1 | tail -2 ) > gpurun_out/${TAG}_bench_c3.log
( timeout 400 python bench.py --config C3 --impl reference --steps 3 --warmup 1 2>&1 | tail -2 ) > gpurun_out/${TAG}_bench_c3_reference.log
# PDL on/off at a launch-bound size (7 audios): does the programmatic edge survive graph capture?
python - > gpurun_out/${TAG}_pdl_small.log 2>&1 <<'PY'
import os, sys, time, tempfile, subprocess
code = r'''
import os, sys, time, tempfile
sys.path.insert(0, os.getcwd())
import numpy as np, bench
from fakebob_b200 import synth
root = tempfile.mkdtemp()
tree = bench.build_gmm_tree_gpu(root, 0, 5)
model = bench.make_model("C2", tree, None, 0, root + "/g")
audio = synth.synth_utterance(0, 0, 80000)
from fakebob_b200.FAKEBOB import FakeBob
for S in (6, 50):
    fb = FakeBob("OSI", "untargeted", model, max_iter=300, samples_per_draw=S, seed=1, verbose=False, iters_per_launch=300)
    fb.attack(audio, None, threshold=1e6)
    t0 = time.perf_counter(); fb.attack(audio, None, threshold=1e6); dt = time.perf_counter() - t0
    print("PDL", os.environ.get("FB_NO_PDL", "on"), "S", S, "us/iter %.1f" % (dt / 300 * 1e6))
'''
for env in ({}, {"FB_NO_PDL": "1"}):
    e = dict(os.environ); e.update(env)
    print(subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True).stdout)
PY
echo done
