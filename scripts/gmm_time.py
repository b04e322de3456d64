"""Times the scoring path stages on the bench batch (51 x 5 s, UBM + 5 MAP-adapted speaker GMMs x 2048) with the per-stage
profiler.  FB_LIB_PATH selects a kernel variant (scripts/build_variant.sh), FAKEBOB_GMM_DELTA_TERMS the contraction."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from fakebob_b200 import synth
from fakebob_b200.engine import GmmEngine, to_audio_list
root = os.environ.get("FB_TREE_ROOT") or tempfile.mkdtemp(prefix="fakebob_time_")
marker = os.path.join(root, "tree.npy")
if os.path.exists(marker):
    paths = list(np.load(marker, allow_pickle=True))
else:
    tree = bench.build_gmm_tree_gpu(root, 0, 5)
    paths = [tree["ubm"]] + [m[2] for m in tree["models"]]
    np.save(marker, np.array(paths, dtype=object), allow_pickle=True)
eng = GmmEngine.from_files(paths)
audio = synth.synth_utterance(0, 0, 80000)
r = np.random.default_rng(1)
big = to_audio_list([audio + 0.001 * r.standard_normal(80000) for i in range(51)])     # a NES batch: perturbations of one audio
eng.score_avg_ll(big)
eng.profile(True)
for _ in range(20):
    eng.score_avg_ll(big)
p = eng.profile_read()
print("lib", os.environ.get("FB_LIB_PATH", "default"), eng.gmm_info(), "rows", eng.voiced_rows(),
      {k: round(v[0] / max(v[1], 1) * 1e3, 1) for k, v in p.items() if v[1]})
