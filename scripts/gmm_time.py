"""Times the scoring path stages on the bench batch (51 x 5 s, 6 GMMs x 2048) with the per-stage profiler."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fakebob_b200 import synth
from fakebob_b200.engine import GmmEngine, to_audio_list
r = np.random.default_rng(0)
C = 2048
gm = []
iv = r.uniform(0.5, 4.0, (C, 72)).astype(np.float32)          # shared by all models (MAP mean-only adaptation)
for m in range(6):
    mu = r.standard_normal((C, 72)).astype(np.float32)
    w = np.full(C, 1.0 / C, np.float32)
    gc = (np.log(w) - 0.5 * (72 * np.log(2 * np.pi) - np.log(iv).sum(1) + (mu * mu * iv).sum(1))).astype(np.float32)
    gm.append({"weights": w, "means_invvars": mu * iv, "inv_vars": iv, "gconsts": gc})
eng = GmmEngine(gm)
big = to_audio_list([synth.synth_utterance(100 + i, i % 5, 80000) for i in range(51)])
eng.score_avg_ll(big)
eng.profile(True)
for _ in range(20):
    eng.score_avg_ll(big)
p = eng.profile_read()
print("rows", eng.voiced_rows(), {k: round(v[0] / max(v[1], 1) * 1e3, 1) for k, v in p.items() if v[1]})
