"""Bring-up diagnostics on a real B200: per-stage max errors vs the oracle, simt vs tcgen05 GMM."""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fakebob_b200 import synth, kaldi_io            # noqa: E402
from fakebob_b200.engine import GmmEngine, to_audio_list   # noqa: E402
from oracle import kaldi_feats as kf                 # noqa: E402
from oracle.diag_gmm import DiagGmm                  # noqa: E402


def main():
    C = int(os.environ.get("PROBE_C", "256"))
    root = tempfile.mkdtemp()
    t0 = time.time()
    tree = synth.build_gmm_tree(root, kf.voiced_features, n_speakers=3, C=C, n_ubm_utts=12, n_samples=32000,
                                n_znorm_utts=4, em_iters=2)
    print("tree built in %.1fs" % (time.time() - t0), flush=True)
    paths = [tree["ubm"]] + [m[2] for m in tree["models"]]
    eng = GmmEngine.from_files(paths)
    eng.set_debug(True)
    gm = []
    for p in paths:
        g = kaldi_io.read_diag_gmm(p)
        gm.append(DiagGmm(g["weights"], g["means_invvars"], g["inv_vars"], g["gconsts"]))
    audios = [synth.synth_utterance(3, 0, 32000), synth.synth_utterance(4, 1, 24000), synth.synth_utterance(5, 2, 40001)]
    lst = to_audio_list(audios)
    for impl in ("simt", "umma"):
        eng.set_gmm_impl(impl)
        try:
            avg = eng.score_avg_ll(lst)
        except Exception as e:
            print(impl, "FAILED:", e, flush=True)
            continue
        st = eng.last_stages()
        f0 = r0 = 0
        for b, w in enumerate(lst):
            m = kf.mfcc(w)
            v = kf.compute_vad(m)
            f = kf.sliding_cmn(kf.add_deltas(m))[v != 0]
            T = m.shape[0]
            gmf = st["mfcc"][f0:f0 + T]
            gv = (st["vad"][f0:f0 + T] >= 0).astype(np.float32)
            Tv = int(st["voiced"][b])
            gf = st["feats"][r0:r0 + Tv]
            print("[%s] utt %d T=%d Tv=%d/%d  mfcc err %.3g  vad diff %d  feat err %s" % (
                impl, b, T, Tv, int(v.sum()), np.abs(gmf - m).max(), int((gv != v).sum()),
                ("%.3g" % np.abs(gf - f).max()) if gf.shape == f.shape else "shape"), flush=True)
            for k, g in enumerate(gm):
                ref = g.frame_loglikes(gf)
                got = st["frame_ll"][k][r0:r0 + Tv]
                print("     model %d frame_ll err %.3g (|ref| %.1f)  avg got %.6f want %.6f" % (
                    k, np.abs(got - ref).max(), np.abs(ref).mean(), avg[b, k], float(g.avg_loglike(gf))), flush=True)
            f0 += T
            r0 += Tv
    # timing of the score path at the bench batch size
    big = to_audio_list([synth.synth_utterance(100 + i, i % 5, 80000) for i in range(51)])
    eng.set_debug(False)
    for impl in ("simt", "umma"):
        eng.set_gmm_impl(impl)
        eng.score_avg_ll(big)
        t0 = time.time()
        for _ in range(5):
            eng.score_avg_ll(big)
        print("[%s] score 51 x 5 s, %d models, C=%d: %.3f ms / call (host buffers)" % (
            impl, len(paths), C, (time.time() - t0) / 5 * 1e3), flush=True)


if __name__ == "__main__":
    main()
