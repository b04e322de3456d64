"""Seeded synthetic audio and synthetic ``pre-models/`` + ``model/`` trees.

The reference's pre-trained models (``pre-models.tgz``, README.md:118) and data
(README.md:84) are not available offline, so benchmarks and tests run on synthetic
stand-ins with the right shapes: a 2048-mixture diagonal UBM
(``final.dubm``), MAP mean-adapted speaker GMMs (``<spk>-identity.gmm`` as written by
``build_spk_models.py:202-219``), a full-covariance UBM (``final.ubm``), an i-vector
extractor (``final.ie``), the LDA/PLDA back-end (``mean.vec``, ``transform.mat``,
``plda``) and the 5-element speaker pickles (``build_spk_models.py:152,272``).

Everything is written in Kaldi's own binary formats (``kaldi_io``), so the scorer
classes load synthetic and real trees through the same code.

``feature_fn(int16 wave) -> (Tv, 72) float32`` is injected by the caller: the bench
passes the CUDA front-end, CPU tests pass the oracle's.  This module itself holds no
scoring arithmetic of the hot path; enrolment (MAP, z-norm) is offline work
 done here in float64 numpy.
"""
import os
import pickle

import numpy as np

from . import kaldi_io

FS = 16000


# ----------------------------------------------------------------------------- audio
def speaker_profile(spk_seed):
    r = np.random.default_rng(10_000 + int(spk_seed))
    return {
        "f0": r.uniform(95.0, 230.0),
        "formant_shift": r.uniform(0.85, 1.2),
        "tilt": r.uniform(0.6, 1.4),
        "breath": r.uniform(0.01, 0.05),
    }


def synth_utterance(seed, spk_seed=0, n_samples=80000, peak=0.3, fs=FS):
    """Speech-like test signal: harmonic source through slowly changing formant gains,
    syllabic envelope with ~25 % low-energy gaps, plus breath noise.  Returned as
    int16-exact float64 in [-1, 1] (like attackMain.py:122-123: wav / 2**15)."""
    prof = speaker_profile(spk_seed)
    r = np.random.default_rng(1234 + 7919 * int(seed) + 104729 * int(spk_seed))
    t = np.arange(n_samples) / fs
    dur = n_samples / fs
    # syllable segmentation
    bounds = [0.0]
    while bounds[-1] < dur:
        bounds.append(bounds[-1] + r.uniform(0.12, 0.38))
    bounds = np.asarray(bounds)
    n_seg = len(bounds) - 1
    voiced = r.uniform(size=n_seg) > 0.36
    level = np.where(voiced, r.uniform(0.35, 1.0, size=n_seg), r.uniform(0.0005, 0.004, size=n_seg))
    centres = 0.5 * (bounds[:-1] + bounds[1:])
    env = np.interp(t, centres, level)
    env *= 0.75 + 0.25 * np.sin(2 * np.pi * r.uniform(3.0, 6.0) * t + r.uniform(0, 6.28))
    # formant tracks per segment
    base_formants = np.array([550.0, 1500.0, 2500.0, 3600.0]) * prof["formant_shift"]
    F = base_formants[None, :] * r.uniform(0.7, 1.35, size=(n_seg, 4))
    bw = np.array([90.0, 130.0, 180.0, 260.0])
    f0 = prof["f0"] * (1.0 + 0.04 * np.sin(2 * np.pi * r.uniform(1.5, 4.0) * t) + 0.02 * np.sin(2 * np.pi * 0.4 * t))
    phase = 2 * np.pi * np.cumsum(f0) / fs
    n_h = int(min(48, 7000.0 / prof["f0"]))
    x = np.zeros(n_samples)
    for h in range(1, n_h + 1):
        fh = h * prof["f0"]
        gain_seg = np.sum(np.exp(-0.5 * ((fh - F) / bw[None, :]) ** 2), axis=1) + 0.02
        gain = np.interp(t, centres, gain_seg) / (h ** prof["tilt"])
        x += gain * np.sin(h * phase + r.uniform(0, 6.28))
    x *= env
    noise = r.standard_normal(n_samples)
    noise = np.convolve(noise, np.array([0.25, 0.5, 0.25]), mode="same")
    x = x / (np.max(np.abs(x)) + 1e-12)
    x = x + prof["breath"] * noise * (0.2 + env) + 2e-4 * r.standard_normal(n_samples)
    x = peak * x / np.max(np.abs(x))
    return np.trunc(x * 32768.0) / 32768.0


def to_int16(audio):
    a = np.asarray(audio)
    if a.dtype == np.int16:
        return a
    return (a * 32768.0).astype(np.int16)


# ----------------------------------------------------------------------------- diag UBM + speakers
def _diag_loglikes(X, w, mu, var):
    iv = 1.0 / var
    gc = np.log(w) - 0.5 * (X.shape[1] * np.log(2 * np.pi) + np.log(var).sum(1) + (mu * mu * iv).sum(1))
    return X @ (mu * iv).T - 0.5 * (X * X) @ iv.T + gc[None, :]


def train_diag_ubm(X, C, seed=7, n_iter=3, var_floor_frac=0.02):
    """A few EM iterations from a random-frame initialisation (float64)."""
    r = np.random.default_rng(seed)
    X = np.asarray(X, dtype=np.float64)
    T, D = X.shape
    gvar = X.var(axis=0) + 1e-6
    idx = r.choice(T, size=C, replace=T < C)
    mu = X[idx] + 0.05 * np.sqrt(gvar)[None, :] * r.standard_normal((C, D))
    var = np.tile(gvar * 0.35, (C, 1)) * r.uniform(0.7, 1.3, size=(C, D))
    w = r.dirichlet(np.full(C, 5.0))
    floor = var_floor_frac * gvar
    for _ in range(n_iter):
        occ = np.zeros(C)
        s1 = np.zeros((C, D))
        s2 = np.zeros((C, D))
        for a in range(0, T, 8192):
            xb = X[a:a + 8192]
            ll = _diag_loglikes(xb, w, mu, var)
            ll -= ll.max(axis=1, keepdims=True)
            p = np.exp(ll)
            p /= p.sum(axis=1, keepdims=True)
            occ += p.sum(0)
            s1 += p.T @ xb
            s2 += p.T @ (xb * xb)
        ok = occ > 1e-3
        mu = np.where(ok[:, None], s1 / np.maximum(occ, 1e-10)[:, None], mu)
        v = s2 / np.maximum(occ, 1e-10)[:, None] - mu * mu
        var = np.where(ok[:, None], np.maximum(v, floor[None, :]), var)
        w = np.maximum(occ, 1e-3)
        w = w / w.sum()
    return w, mu, var


def map_adapt_means(X, w, mu, var, tau=10.0):
    """gmm-global-acc-stats + gmm-global-est-map --update-flags=m (build_spk_models.py:202-219)."""
    X = np.asarray(X, dtype=np.float64)
    ll = _diag_loglikes(X, w, mu, var)
    ll -= ll.max(axis=1, keepdims=True)
    p = np.exp(ll)
    p /= p.sum(axis=1, keepdims=True)
    occ = p.sum(0)
    return (p.T @ X + tau * mu) / (occ + tau)[:, None]


def _avg_ll(X, w, mu, var):
    ll = _diag_loglikes(np.asarray(X, dtype=np.float64), w, mu, var)
    m = ll.max(axis=1)
    return float(np.mean(m + np.log(np.exp(ll - m[:, None]).sum(axis=1))))


def write_conf(pre_dir):
    os.makedirs(os.path.join(pre_dir, "conf"), exist_ok=True)
    with open(os.path.join(pre_dir, "conf", "mfcc.conf"), "w") as f:
        f.write("--sample-frequency=16000\n--frame-length=25 # the default is 25\n--low-freq=20 # the default.\n"
                "--high-freq=7600 # the default is zero meaning use the Nyquist (8k in this case).\n"
                "--num-mel-bins=30\n--num-ceps=24\n--snip-edges=false\n")
    with open(os.path.join(pre_dir, "conf", "vad.conf"), "w") as f:
        f.write("--vad-energy-threshold=5.5\n--vad-energy-mean-scale=0.5\n"
                "--vad-proportion-threshold=0.12\n--vad-frames-context=2\n")
    with open(os.path.join(pre_dir, "delta_opts"), "w") as f:
        f.write("--delta-window=3 --delta-order=2\n")


def build_gmm_tree(root, feature_fn, n_speakers=5, C=2048, n_ubm_utts=64, n_samples=80000,
                   seed=7, n_znorm_utts=8, em_iters=3):
    """Create ``<root>/pre-models`` (final.dubm + conf) and ``<root>/model`` (identity GMMs + pickles).

    Returns dict(pre_model_dir, model_dir, ubm, spk_ids, models=[5-lists], ubm_params)."""
    pre_dir = os.path.join(root, "pre-models")
    model_dir = os.path.join(root, "model")
    os.makedirs(pre_dir, exist_ok=True)
    os.makedirs(model_dir, exist_ok=True)
    write_conf(pre_dir)
    feats = []
    for u in range(n_ubm_utts):
        a = synth_utterance(seed=1000 + u, spk_seed=100 + (u % 16), n_samples=n_samples)
        feats.append(np.asarray(feature_fn(to_int16(a)), dtype=np.float64))
    X = np.concatenate(feats, axis=0)
    w, mu, var = train_diag_ubm(X, C, seed=seed, n_iter=em_iters)
    iv = 1.0 / var
    ubm_path = os.path.join(pre_dir, "final.dubm")
    kaldi_io.write_diag_gmm(ubm_path, w, mu * iv, iv)
    # z-norm cohort (only used by gmm_CSI, build_spk_models.py:259-261)
    zfeats = [np.asarray(feature_fn(to_int16(synth_utterance(seed=5000 + u, spk_seed=200 + u, n_samples=n_samples))),
                         dtype=np.float64) for u in range(n_znorm_utts)]
    spk_ids, models = [], []
    for s in range(n_speakers):
        spk_id = "%04d" % (1580 + 37 * s)
        enrol = synth_utterance(seed=9000 + s, spk_seed=s, n_samples=n_samples)
        Xe = np.asarray(feature_fn(to_int16(enrol)), dtype=np.float64)
        mu_s = map_adapt_means(Xe, w, mu, var, tau=10.0)
        ident = os.path.abspath(os.path.join(model_dir, spk_id + "-identity.gmm"))
        kaldi_io.write_diag_gmm(ident, w, mu_s * iv, iv)
        zs = np.array([_avg_ll(zf, w, mu_s, var) for zf in zfeats])
        model = [spk_id, spk_id + "-enroll", ident, float(zs.mean()), float(zs.std() + 1e-6)]
        with open(os.path.join(model_dir, spk_id + ".gmm"), "wb") as f:
            pickle.dump(model, f, protocol=-1)
        spk_ids.append(spk_id)
        models.append(model)
    return {"pre_model_dir": pre_dir, "model_dir": model_dir, "ubm": ubm_path, "spk_ids": spk_ids,
            "models": models, "ubm_params": (w, mu, var)}


# ----------------------------------------------------------------------------- i-vector / PLDA side
def build_ivector_params(root, ubm_params, R=400, L=200, seed=11, rank=4):
    """Write final.ubm, final.ie, mean.vec, transform.mat, plda under <root>/pre-models.

    Full UBM: diag UBM means/weights, covariance = diag + low-rank SPD perturbation.
    Extractor: prior_offset=100, M_c[:,0] = mu_c/100, M_c[:,1:] ~ N(0, 0.05^2)*sigma_c,
    Sigma_c^-1 from the full UBM, no weight projection."""
    pre_dir = os.path.join(root, "pre-models")
    os.makedirs(pre_dir, exist_ok=True)
    w, mu, var = ubm_params
    C, D = mu.shape
    r = np.random.default_rng(seed)
    inv_covars = np.empty((C, D, D), dtype=np.float64)
    means_invcovars = np.empty((C, D), dtype=np.float64)
    gconsts = np.empty(C, dtype=np.float64)
    sd = np.sqrt(var)
    for c in range(C):
        U = r.standard_normal((D, rank)) * (0.35 * sd[c])[:, None]
        cov = np.diag(var[c]) + U @ U.T
        ic = np.linalg.inv(cov)
        ic = 0.5 * (ic + ic.T)
        inv_covars[c] = ic
        means_invcovars[c] = ic @ mu[c]
        _, logdet = np.linalg.slogdet(cov)
        gconsts[c] = np.log(w[c]) - 0.5 * (D * np.log(2 * np.pi) + logdet + mu[c] @ ic @ mu[c])
    kaldi_io.write_full_gmm(os.path.join(pre_dir, "final.ubm"), w, means_invcovars, inv_covars, gconsts)
    prior_offset = 100.0
    M = r.standard_normal((C, D, R)) * 0.3 * sd[:, :, None]
    M[:, :, 0] = mu / prior_offset
    # Kaldi stores Sigma_inv as float64 SpMatrix; derive from the float32-rounded full UBM like a real run
    sig_inv = inv_covars.astype(np.float32).astype(np.float64)
    kaldi_io.write_ivector_extractor(os.path.join(pre_dir, "final.ie"), w, M, sig_inv, prior_offset)
    mean_vec = r.standard_normal(R) * 0.3
    mean_vec[0] += 0.0
    kaldi_io.write_vector(os.path.join(pre_dir, "mean.vec"), mean_vec)
    lda = r.standard_normal((L, R)) / np.sqrt(R)
    lda_off = r.standard_normal((L, 1)) * 0.05
    kaldi_io.write_matrix(os.path.join(pre_dir, "transform.mat"), np.concatenate([lda, lda_off], axis=1))
    q, _ = np.linalg.qr(r.standard_normal((L, L)))
    transform = q * r.uniform(0.6, 1.6, size=(1, L))
    psi = np.sort(r.uniform(0.05, 12.0, size=L))[::-1].copy()
    kaldi_io.write_plda(os.path.join(pre_dir, "plda"), r.standard_normal(L) * 0.1, transform, psi)
    return pre_dir


def build_ivector_speakers(root, ivector_fn, plda_score_fn, n_speakers=5, n_samples=80000, n_znorm_utts=8):
    """Enrol synthetic speakers: one i-vector each, written as a Kaldi text ark whose scp target is the
    pickle's identity_location (build_spk_models.py:141-152).  z-norm from an impostor cohort (:124-131).

    ivector_fn(int16 wave) -> (R,) float64 raw i-vector;  plda_score_fn(enrolled (K,R), test (B,R)) -> (B,K)."""
    model_dir = os.path.join(root, "model")
    iv_dir = os.path.join(root, "enroll-ivectors")
    os.makedirs(model_dir, exist_ok=True)
    os.makedirs(iv_dir, exist_ok=True)
    spk_ids, utts, ivs = [], [], []
    for s in range(n_speakers):
        spk_id = "%04d" % (1580 + 37 * s)
        enrol = synth_utterance(seed=9000 + s, spk_seed=s, n_samples=n_samples)
        ivs.append(np.asarray(ivector_fn(to_int16(enrol)), dtype=np.float64))
        spk_ids.append(spk_id)
        utts.append(spk_id + "-enroll")
    # text round trip ('ark,t', 7 significant digits) exactly like the reference's enrolled identities
    targets = kaldi_io.write_text_vector_ark(os.path.abspath(os.path.join(iv_dir, "ivector.1.ark")), list(zip(utts, ivs)))
    enrolled = np.stack([kaldi_io.read_vector(targets[u]) for u in utts])
    cohort = np.stack([np.asarray(ivector_fn(to_int16(synth_utterance(seed=5000 + u, spk_seed=200 + u, n_samples=n_samples))),
                                  dtype=np.float64) for u in range(n_znorm_utts)])
    zs = np.asarray(plda_score_fn(enrolled, cohort))          # (n_cohort, K)
    models = []
    for k, (spk_id, utt) in enumerate(zip(spk_ids, utts)):
        model = [spk_id, utt, targets[utt], float(zs[:, k].mean()), float(zs[:, k].std() + 1e-6)]
        with open(os.path.join(model_dir, spk_id + ".iv"), "wb") as f:
            pickle.dump(model, f, protocol=-1)
        models.append(model)
    return {"spk_ids": spk_ids, "models": models, "enrolled": enrolled}
