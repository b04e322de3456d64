"""Generates tests/golden/kaldi_stages.npz: per-stage outputs of the ORACLE (oracle/kaldi_feats.py, oracle/diag_gmm.py) on a
seeded synthetic utterance and a seeded 128-component diagonal GMM stored alongside.

No Kaldi exists in this image (parity with Kaldi is unpinned), so this fixture does not pin the oracle
to the reference; it freezes the oracle's arithmetic -- a change in oracle/ or in the CUDA path shows up against a committed
artefact, not only against the oracle's current code.   python tests/golden/make_golden_kaldi.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from fakebob_b200 import synth  # noqa: E402
from oracle import kaldi_feats as kf  # noqa: E402
from oracle.diag_gmm import DiagGmm  # noqa: E402


def main():
    audio = synth.synth_utterance(seed=4246, spk_seed=3, n_samples=24000)      # 150 frames, 105 voiced
    wave = kf.float_to_int16(audio)
    mfcc = kf.mfcc(wave)
    vad = kf.compute_vad(mfcc)
    feats = kf.sliding_cmn(kf.add_deltas(mfcc))
    voiced = feats[vad != 0]
    # a GMM that sits on the data: means drawn from the voiced frames, per-dimension variances of the data
    r = np.random.default_rng(99)
    C = 128
    mu = voiced[r.integers(0, voiced.shape[0], C)].astype(np.float64) + 0.3 * r.standard_normal((C, 72))
    var = voiced.astype(np.float64).var(axis=0)[None, :] * r.uniform(0.5, 1.5, (C, 72))
    w = r.dirichlet(np.full(C, 5.0))
    ubm = DiagGmm.from_moments(w, mu, var)
    spk = ubm.map_adapt_means(voiced[::2], tau=10.0)
    spk_all = ubm.map_adapt_means(voiced, tau=10.0)      # what enrolling this very utterance gives (fb_map_adapt_host)
    out = os.path.join(HERE, "kaldi_stages.npz")
    np.savez_compressed(
        out, wave=wave, mfcc=mfcc.astype(np.float32), vad=vad.astype(np.int8), voiced_feats=voiced.astype(np.float32),
        ubm_weights=ubm.weights, ubm_means_invvars=ubm.means_invvars, ubm_inv_vars=ubm.inv_vars, ubm_gconsts=ubm.gconsts,
        spk_means_invvars=spk.means_invvars, spk_gconsts=spk.gconsts,
        spk_all_means_invvars=spk_all.means_invvars, spk_all_gconsts=spk_all.gconsts, spk_all_occupancy=spk_all.occupancy,
        frame_ll_ubm=ubm.frame_loglikes(voiced), frame_ll_spk=spk.frame_loglikes(voiced),
        avg_ll=np.array([float(ubm.avg_loglike(voiced)), float(spk.avg_loglike(voiced))]))
    print(out, "frames", mfcc.shape[0], "voiced", voiced.shape[0], "avg ll", float(ubm.avg_loglike(voiced)), float(spk.avg_loglike(voiced)))


if __name__ == "__main__":
    main()
