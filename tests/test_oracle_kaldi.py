"""Guards the Kaldi restatement (oracle/kaldi_feats.py, oracle/diag_gmm.py) against independent implementations
available in-container.  These do not pin parity with real Kaldi (absent here) -- see oracle/__init__.py."""
import numpy as np
import pytest

from fakebob_b200 import synth
from oracle import kaldi_feats as kf
from oracle.diag_gmm import DiagGmm, log_sum_exp_rows


def test_mfcc_matches_torchaudio_kaldi_port():
    torch = pytest.importorskip("torch")
    ta = pytest.importorskip("torchaudio")
    w = synth.to_int16(synth.synth_utterance(5, 2, 32000))
    m = kf.mfcc(w)
    ref = ta.compliance.kaldi.mfcc(torch.from_numpy(w.astype(np.float32))[None], dither=0.0, energy_floor=0.0, use_energy=True,
                                   num_mel_bins=30, num_ceps=24, low_freq=20, high_freq=7600, snip_edges=False,
                                   sample_frequency=16000.0, raw_energy=True).numpy()
    assert m.shape == ref.shape == (200, 24)
    assert np.abs(m - ref).max() < 2e-3


def test_frame_count_and_edges():
    cfg = kf.FeatConfig()
    assert kf.num_frames(80000, cfg) == 500 and kf.num_frames(80079, cfg) == 500 and kf.num_frames(80080, cfg) == 501
    fr = kf.extract_frames(np.arange(1000, dtype=np.int16), cfg)
    assert fr.shape == (6, 400)
    assert fr[0, 0] == 119 and fr[0, 119] == 0 and fr[0, 120] == 0 and fr[0, 121] == 1      # reflected left edge
    assert fr[-1, -1] == 2 * 1000 - 1 - (160 * 5 - 120 + 399)


def test_delta_scales_and_clamping():
    sc = kf.delta_scales(kf.FeatConfig())
    assert np.allclose(sc[1], np.arange(-3, 4) / 28.0)
    assert len(sc[2]) == 13 and abs(sc[2].sum()) < 1e-6
    x = np.arange(20, dtype=np.float32)[:, None] * np.ones((1, 2), np.float32)
    d = kf.add_deltas(x)
    assert d.shape == (20, 6)
    assert np.allclose(d[5:15, 2:4], 1.0, atol=1e-5) and np.allclose(d[6:14, 4:6], 0.0, atol=1e-5)
    assert d[0, 2] < 1.0                                            # clamped edge


def test_sliding_cmn_windows():
    r = np.random.default_rng(0)
    x = r.standard_normal((500, 3)).astype(np.float32)
    y = kf.sliding_cmn(x)
    assert np.allclose(y[0], x[0] - x[:300].mean(0), atol=1e-5)
    assert np.allclose(y[150], x[150] - x[:300].mean(0), atol=1e-5)
    assert np.allclose(y[250], x[250] - x[100:400].mean(0), atol=1e-5)
    assert np.allclose(y[499], x[499] - x[200:].mean(0), atol=1e-5)
    z = kf.sliding_cmn(x[:120])
    assert np.allclose(z, x[:120] - x[:120].mean(0), atol=1e-5)


def test_vad_rule():
    m = np.zeros((10, 24), np.float32)
    m[:, 0] = [0, 0, 0, 0, 20, 0, 0, 0, 0, 0]
    v = kf.compute_vad(m)            # threshold 5.5 + 0.5 * 2 = 6.5; frame 4 above; +-2 context
    assert v.tolist() == [0, 0, 1, 1, 1, 1, 1, 0, 0, 0]


def test_diag_gmm_matches_sklearn():
    sk = pytest.importorskip("sklearn.mixture")
    r = np.random.default_rng(1)
    C, D, T = 32, 72, 300
    w = r.dirichlet(np.full(C, 5.0))
    mu = r.standard_normal((C, D)) * 2
    var = r.uniform(0.3, 2.0, (C, D))
    X = (mu[r.integers(0, C, T)] + r.standard_normal((T, D))).astype(np.float32)
    g = DiagGmm.from_moments(w, mu, var)
    gm = sk.GaussianMixture(n_components=C, covariance_type="diag")
    gm.weights_, gm.means_, gm.covariances_ = w, mu, var
    gm.precisions_cholesky_ = 1.0 / np.sqrt(var)
    ref = gm.score_samples(X.astype(np.float64))
    got = g.frame_loglikes(X)
    assert np.abs(got - ref).max() < 2e-3 * np.abs(ref).max() / 100 + 2e-3
    assert abs(float(g.avg_loglike(X)) - ref.mean()) < 1e-3
    post = g.posteriors(X)
    assert np.abs(post - gm.predict_proba(X.astype(np.float64))).max() < 1e-4


def test_log_sum_exp_prunes_like_kaldi():
    ll = np.array([[0.0, -10.0, -15.9, -16.0, -100.0]], dtype=np.float32)
    want = np.log(1.0 + np.exp(np.float32(-10.0)) + np.exp(np.float32(-15.9)))       # -16 and -100 are below the cutoff
    assert abs(float(log_sum_exp_rows(ll)[0]) - want) < 1e-7


def test_map_adaptation_moves_means_toward_data():
    r = np.random.default_rng(2)
    C, D = 8, 72
    g = DiagGmm.from_moments(np.full(C, 1 / C), r.standard_normal((C, D)), np.ones((C, D)))
    X = (g.means()[3] + 0.5 + 0.1 * r.standard_normal((200, D))).astype(np.float32)
    a = g.map_adapt_means(X, tau=10.0)
    assert np.allclose(a.inv_vars, g.inv_vars) and np.allclose(a.weights, g.weights)
    assert np.linalg.norm(a.means()[3] - X.mean(0)) < np.linalg.norm(g.means()[3] - X.mean(0)) * 0.2


def test_int16_truncation_toward_zero():
    a = np.array([0.99999, -0.99999, 1.4 / 32768, -1.6 / 32768])
    assert kf.float_to_int16(a).tolist() == [32767, -32767, 1, -1]


def test_oracle_reproduces_committed_stage_fixture():
    """tests/golden/kaldi_stages.npz (made by tests/golden/make_golden_kaldi.py) freezes the oracle's per-stage arithmetic."""
    import os
    from oracle import kaldi_feats as kf
    from oracle.diag_gmm import DiagGmm
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kaldi_stages.npz"))
    wave = g["wave"]
    m = kf.mfcc(wave)
    v = kf.compute_vad(m)
    f = kf.sliding_cmn(kf.add_deltas(m))[v != 0]
    assert m.shape == g["mfcc"].shape and np.abs(m - g["mfcc"]).max() < 2e-4        # BLAS / libm builds may differ in the last bits
    assert np.array_equal(v != 0, g["vad"] != 0) and 0 < int((v != 0).sum()) < m.shape[0]
    assert np.abs(f - g["voiced_feats"]).max() < 2e-4
    ubm = DiagGmm(g["ubm_weights"], g["ubm_means_invvars"], g["ubm_inv_vars"], g["ubm_gconsts"])
    spk = DiagGmm(g["ubm_weights"], g["spk_means_invvars"], g["ubm_inv_vars"], g["spk_gconsts"])
    X = g["voiced_feats"]
    assert np.abs(ubm.frame_loglikes(X) - g["frame_ll_ubm"]).max() < 2e-4
    assert np.abs(spk.frame_loglikes(X) - g["frame_ll_spk"]).max() < 2e-4
    assert abs(float(ubm.avg_loglike(X)) - g["avg_ll"][0]) < 1e-4 and abs(float(spk.avg_loglike(X)) - g["avg_ll"][1]) < 1e-4
    re = ubm.map_adapt_means(X[::2], tau=10.0)
    assert np.abs(re.means_invvars - g["spk_means_invvars"]).max() < 2e-4 and np.abs(re.gconsts - g["spk_gconsts"]).max() < 2e-4
    re = ubm.map_adapt_means(X, tau=10.0)
    assert np.abs(re.means_invvars - g["spk_all_means_invvars"]).max() < 2e-4 and np.abs(re.occupancy - g["spk_all_occupancy"]).max() < 1e-3


def test_oracle_reproduces_committed_fullsize_fixture(tmp_path):
    """tests/golden/fullsize_c2.npz (tests/golden/make_golden_fullsize.py) at the benchmark's model size: the seeded
    2048-mixture tree is rebuilt and the oracle re-run on the stored utterance 0 (the C3 half needs 6 GB and minutes: it is
    re-checked by running the generator, not here)."""
    import os
    import fullsize_util as fu
    from oracle import kaldi_feats as kf
    from oracle.diag_gmm import DiagGmm
    from oracle import kaldi_files
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fullsize_c2.npz"))
    tree = fu.build_c2_tree(str(tmp_path))
    assert fu.checksums_close(fu.gmm_checksums(tree), g["checksums"], rtol=1e-5)
    assert np.array_equal(fu.test_wave(0), g["wave0"])
    X = kf.voiced_features(g["wave0"])
    assert X.shape == g["feats0"].shape and np.abs(X - g["feats0"]).max() < 2e-4
    for k, path in enumerate([tree["ubm"]] + [m[2] for m in tree["models"]]):
        p = kaldi_files.read_diag_gmm(path)
        gm = DiagGmm(p["weights"], p["means_invvars"], p["inv_vars"], p["gconsts"])
        assert np.abs(gm.frame_loglikes(X) - g["frame_ll0"][k]).max() < 5e-4
        assert abs(float(gm.avg_loglike(X)) - g["avg_ll0"][k]) < 1e-4
