"""Generates tests/golden/fullsize_c2.npz and tests/golden/fullsize_c3.npz: ORACLE outputs at the sizes bench.py measures.

  C2 (BASELINE.json configs[1], also the model set of configs[3]): 2048-mixture diagonal UBM + 5 MAP-adapted speaker
     GMMs, 5 s utterances -- MFCC / VAD / features / per-frame and average log-likelihoods / OSI, CSI and SV scores for
     three utterances, and a 3-iteration Philox NES trajectory (samples_per_draw = 50) of the oracle attack loop.
  C3 (configs[2], also the extractor of configs[4]): full-covariance 2048-mixture UBM, 400-dim i-vector extractor,
     LDA 200 + PLDA -- Gaussian selection, pruned posteriors, gamma / X statistics, lin, quad, raw i-vector, PLDA
     log-likelihood ratios and z-normed scores for three utterances against three enrolled speakers.

The models are NOT stored (the C3 extractor alone is 0.5 GB): tests/fullsize_util.py rebuilds the same seeded synthetic
trees on the test machine (fakebob_b200/synth.py is deterministic) and checks a few parameter checksums stored here.
The test utterances ARE stored (int16), so a 1-ulp libm difference between hosts cannot move a sample.

Like tests/golden/kaldi_stages.npz this freezes the oracle's arithmetic at the benchmarked sizes; it does not pin the oracle
to Kaldi.  Runs in the CPU container in a few minutes (needs ~6 GB of RAM for the oracle's dense U):

    python tests/golden/make_golden_fullsize.py [c2] [c3]
"""
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fullsize_util as fu  # noqa: E402
from fakebob_b200 import kaldi_io, synth  # noqa: E402
from oracle import kaldi_feats as kf  # noqa: E402
from oracle.diag_gmm import DiagGmm  # noqa: E402


def load_gmm(path):
    from oracle import kaldi_files
    g = kaldi_files.read_diag_gmm(path)
    return DiagGmm(g["weights"], g["means_invvars"], g["inv_vars"], g["gconsts"])


def make_c2(root):
    from oracle.nes import OracleFakeBob, PhiloxNoise
    from oracle.scorers import OracleGmmOSI
    t0 = time.time()
    tree = fu.build_c2_tree(root)
    gmms = [load_gmm(tree["ubm"])] + [load_gmm(m[2]) for m in tree["models"]]
    out = {"checksums": fu.gmm_checksums(tree), "z_means": np.array([m[3] for m in tree["models"]]),
           "z_stds": np.array([m[4] for m in tree["models"]])}
    for u in range(fu.N_UTTS):
        wave = fu.test_wave(u)
        mf = kf.mfcc(wave)
        vad = kf.compute_vad(mf)
        feats = kf.sliding_cmn(kf.add_deltas(mf))
        X = feats[vad != 0]
        fl = np.stack([g.frame_loglikes(X) for g in gmms])
        avg = np.array([float(g.avg_loglike(X)) for g in gmms])
        out["wave%d" % u] = wave
        out["mfcc%d" % u] = mf.astype(np.float32)
        out["vad%d" % u] = vad.astype(np.int8)
        if u == 0:
            out["feats0"] = X.astype(np.float32)
        out["frame_ll%d" % u] = fl.astype(np.float32)
        out["avg_ll%d" % u] = avg
        print("c2 utt %d: %d frames, %d voiced, avg ll %s" % (u, mf.shape[0], X.shape[0], np.round(avg, 4)))
    # NES trajectory: oracle loop + oracle scorer, Philox stream (what FakeBob(rng='philox', seed=...) draws on the device)
    model = OracleGmmOSI(gmms[0], gmms[1:])
    audio = fu.test_wave(0).astype(np.float64) / 32768.0
    fb = OracleFakeBob("OSI", "untargeted", model, max_iter=fu.NES_ITERS, samples_per_draw=fu.NES_S, epsilon=0.002,
                       noise_fn=PhiloxNoise(fu.NES_SEED))
    adv, flag = fb.attack(audio, None, threshold=fu.NES_THRESHOLD)
    out["nes_adver_delta"] = (adv[:, 0].astype(np.int32) - fu.test_wave(0).astype(np.int32)).astype(np.int16)
    out["nes_flag"] = np.array([flag])
    out["nes_distance"] = np.array([float(r[0]) for r in fb.log])
    out["nes_adver_loss"] = np.array([float(np.asarray(r[1]).reshape(-1)[0]) for r in fb.log])
    out["nes_scores"] = np.stack([np.asarray(r[2], dtype=np.float64).reshape(-1) for r in fb.log])
    path = os.path.join(HERE, "fullsize_c2.npz")
    np.savez_compressed(path, **out)
    print(path, "%.0f s, %.2f MB" % (time.time() - t0, os.path.getsize(path) / 1e6))
    return tree


def make_c3(root, tree):
    from oracle.ivector import load_system
    t0 = time.time()
    fu.build_c3_params(root, tree)
    system = load_system(tree["pre_model_dir"])
    print("c3 oracle system loaded (dense U) in %.0f s" % (time.time() - t0))
    spk = synth.build_ivector_speakers(root, system.extract, system.plda_scores, n_speakers=fu.C3_SPEAKERS,
                                       n_samples=fu.N_SAMPLES, n_znorm_utts=fu.C3_ZNORM)
    models = sorted(spk["models"], key=lambda m: m[0])
    enrolled = np.stack([np.asarray(kaldi_io.read_vector(m[2])) for m in models])
    out = {"checksums": fu.iv_checksums(tree["pre_model_dir"]), "enrolled": enrolled,
           "spk_ids": np.array([m[0] for m in models]), "z_means": np.array([m[3] for m in models]),
           "z_stds": np.array([m[4] for m in models])}
    ex = system.extractor
    R = ex.ivector_dim
    for u in range(fu.N_UTTS):
        wave = fu.test_wave(u)
        X = kf.voiced_features(wave)
        gsel, post = system.posteriors(X)
        gamma, Xs = system.stats(X, gsel, post)
        lin = np.einsum("cdr,cd->r", ex.sigma_inv_M, Xs)
        quad = np.einsum("c,crs->rs", gamma, ex.U)
        iv = ex.extract(gamma, Xs).astype(np.float32)
        llr = system.plda_scores(enrolled, iv[None, :])[0]
        out["gsel%d" % u] = gsel.astype(np.int16)
        out["post%d" % u] = post.astype(np.float32)
        out["gamma%d" % u] = gamma
        out["xsum%d" % u] = Xs.sum(axis=1)                       # per-component sum over the 72 dims (full X only for utt 0)
        if u == 0:
            out["X0"] = Xs.astype(np.float32)
            out["quad0_tril"] = quad[np.tril_indices(R)].astype(np.float32)
        out["lin%d" % u] = lin
        out["quad_diag%d" % u] = np.diag(quad).copy()
        out["quad_row0_%d" % u] = quad[0].copy()
        out["ivector%d" % u] = iv
        out["llr%d" % u] = llr
        out["score%d" % u] = (llr - out["z_means"]) / out["z_stds"]
        print("c3 utt %d: %d voiced, active comps %d, |iv| %.3f, llr %s" % (u, X.shape[0], int((gamma != 0).sum()),
                                                                          float(np.linalg.norm(iv)), np.round(llr, 3)))
    path = os.path.join(HERE, "fullsize_c3.npz")
    np.savez_compressed(path, **out)
    print(path, "%.0f s, %.2f MB" % (time.time() - t0, os.path.getsize(path) / 1e6))


def main():
    what = set(sys.argv[1:]) or {"c2", "c3"}
    root = tempfile.mkdtemp(prefix="fakebob_fullsize_")
    if "c2" in what:
        tree = make_c2(root)
    else:
        tree = fu.build_c2_tree(root)
    if "c3" in what:
        make_c3(root, tree)


if __name__ == "__main__":
    main()
