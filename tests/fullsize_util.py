"""Shared by tests/golden/make_golden_fullsize.py (fixture generator, CPU container) and tests/test_gpu_fullsize.py
(GPU box): the seeded synthetic model trees at BASELINE.json's sizes and the test utterances.

The trees are rebuilt on the test machine because they cannot be committed (the R = 400 extractor is 0.5 GB); the
generator stores parameter checksums so a tree that came out differently is reported instead of silently compared."""
import os

import numpy as np

C2_MIX = 2048
C2_SPEAKERS = 5
N_SAMPLES = 80000
N_UTTS = 3
NES_S, NES_ITERS, NES_SEED, NES_THRESHOLD = 50, 3, 20261017, 1e3
C3_R, C3_L, C3_SPEAKERS, C3_ZNORM = 400, 200, 3, 4


def test_wave(u):
    """int16 test utterance u (5 s); the fixtures store these, the generator and a fresh machine agree unless libm differs."""
    from fakebob_b200 import synth
    return synth.to_int16(synth.synth_utterance(seed=700 + u, spk_seed=u % 3, n_samples=N_SAMPLES))


def build_c2_tree(root):
    """pre-models/final.dubm + model/<spk>-identity.gmm exactly as bench.py's CPU arm builds them (oracle features)."""
    from fakebob_b200 import synth
    from oracle import kaldi_feats as kf
    tree = synth.build_gmm_tree(root, kf.voiced_features, n_speakers=C2_SPEAKERS, C=C2_MIX, n_ubm_utts=64, n_samples=N_SAMPLES)
    tree["root"] = root
    return tree


def gmm_checksums(tree):
    from fakebob_b200 import kaldi_io
    out = []
    for p in [tree["ubm"]] + [m[2] for m in tree["models"]]:
        g = kaldi_io.read_diag_gmm(p)
        out.append([float(np.abs(g["means_invvars"].astype(np.float64)).sum()), float(g["inv_vars"].astype(np.float64).sum()),
                    float(g["gconsts"].astype(np.float64).sum())])
    return np.array(out)


def build_c3_params(root, tree):
    """final.ubm / final.ie / mean.vec / transform.mat / plda for R = 400, L = 200 over the C2 UBM (0.5 GB on disk)."""
    from fakebob_b200 import synth
    return synth.build_ivector_params(root, tree["ubm_params"], R=C3_R, L=C3_L)


def iv_checksums(pre_model_dir):
    from fakebob_b200 import kaldi_io
    fg = kaldi_io.read_full_gmm(os.path.join(pre_model_dir, "final.ubm"))
    ie = kaldi_io.read_ivector_extractor(os.path.join(pre_model_dir, "final.ie"))
    pl = kaldi_io.read_plda(os.path.join(pre_model_dir, "plda"))
    return np.array([float(np.abs(fg["inv_covars"].astype(np.float64)).sum()), float(fg["gconsts"].astype(np.float64).sum()),
                     float(np.abs(ie["M"]).sum()), float(np.abs(ie["sigma_inv"]).sum()), float(pl["psi"].sum()),
                     float(np.abs(pl["transform"]).sum())])


def checksums_close(a, b, rtol=1e-6):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b)) + 1e-12))
