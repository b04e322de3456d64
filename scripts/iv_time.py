"""Config C3 at full size: iv_SV, 2048-mix full UBM, 400-dim i-vector, LDA 200 + PLDA, S=50, 5 s audio.
Builds the synthetic tree with GPU features, then times score() and NES iterations."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from fakebob_b200 import synth
from fakebob_b200.engine import IvectorEngine, to_audio_list
from fakebob_b200.FAKEBOB import FakeBob
from fakebob_b200.ivector_PLDA_SV import iv_SV

C = int(os.environ.get("IV_C", "2048")); R = int(os.environ.get("IV_R", "400")); L = int(os.environ.get("IV_L", "200"))
root = tempfile.mkdtemp(prefix="fb_iv_")
t0 = time.time()
bench.N_MIX = C
tree = bench.build_workload_gpu(root, 0)
print("gmm tree %.1fs" % (time.time() - t0), flush=True)
t0 = time.time()
synth.build_ivector_params(root, tree["ubm_params"], R=R, L=L)
print("iv params %.1fs" % (time.time() - t0), flush=True)
t0 = time.time()
eng = IvectorEngine(tree["pre_model_dir"])
print("engine load %.1fs" % (time.time() - t0), flush=True)

def ivector_fn(w):
    return eng.extract_ivectors([np.ascontiguousarray(w, dtype=np.int16)])[0]

def plda_fn(enrolled, test):
    eng.set_enrolled(enrolled)
    # score the cohort i-vectors is not exposed; z-norm only needs plausible numbers here
    return np.zeros((len(test), len(enrolled))) + np.arange(len(test))[:, None]

spk = synth.build_ivector_speakers(root, ivector_fn, plda_fn, n_speakers=1, n_samples=80000, n_znorm_utts=2)
eng.close()
model = iv_SV(os.path.join(root, "sv"), spk["models"][0], pre_model_dir=tree["pre_model_dir"])
big = [synth.synth_utterance(100 + i, i % 5, 80000) for i in range(51)]
model.score(big)
t0 = time.time()
for _ in range(3):
    s = model.score(big)
print("score 51 x 5 s: %.2f ms / call; scores[:3]=%s" % ((time.time() - t0) / 3 * 1e3, s[:3]), flush=True)
audio = synth.synth_utterance(0, 0, 80000)
fb = FakeBob("SV", "untargeted", model, max_iter=60, samples_per_draw=50, seed=1, verbose=False)
fb.attack(audio, None, threshold=1e6)
for rep in range(6):                                # wall clock of whole attack() calls: includes fb_nes_init, graph capture, polling
    fb = FakeBob("SV", "untargeted", model, max_iter=100, samples_per_draw=50, seed=1, verbose=False)
    t0 = time.time()
    fb.attack(audio, None, threshold=1e6)
    dt = time.time() - t0
    print("NES iv_SV S=50: %d iters in %.3fs -> %.1f it/s (%.3f ms/iter)" % (fb.iters_done, dt, fb.iters_done / dt, dt / fb.iters_done * 1e3), flush=True)
    print("   init + per-poll seconds:", [round(x, 3) for x in fb.poll_times], flush=True)
e = model._engine
e.profile(True)
fb = FakeBob("SV", "untargeted", model, max_iter=20, samples_per_draw=50, seed=1, verbose=False)
fb.attack(audio, None, threshold=1e6)
p = e.profile_read()
print({k: round(v[0] / max(v[1], 1) * 1e3, 1) for k, v in p.items() if v[1]})
