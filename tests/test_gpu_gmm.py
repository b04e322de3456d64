"""GPU parity tests for the GMM-UBM scoring path, through the C-ABI (fb_score_gmm_host) and the
reference-named Python wrappers.  Oracle = oracle/ (numpy restatement of the Kaldi arithmetic)."""
import numpy as np
import pytest

from conftest import test_audio as make_audio

pytestmark = pytest.mark.gpu

# float tolerances (float32 pipelines with different summation orders / FFT algorithms)
TOL_MFCC = 2e-3        # absolute, on cepstra of magnitude <= ~60 (log of near-zero mel energies amplifies rounding)
TOL_FEAT = 2e-3
TOL_FRAME_LL = 2e-3    # absolute, per-frame log-likelihood of magnitude ~100
TOL_AVG_LL = 5e-4      # absolute, per-utterance average log-likelihood
TOL_SCORE = 5e-4       # absolute, LLR score (difference of two averages)
# Default (automatic) number of difference terms: speaker slots are scored as slot 0 + x.(w_m - w_0) with ONE fp16 product
# when the MAP offsets are small (fb_set_gmm_delta_terms).  Measured at the benchmark size (scripts/gmm_precision_study.py):
# per-frame deviation <= 6e-3 (7e-4 rms), utterance score deviation <= 1.2e-4 -- the resolution of the reference's own
# 7-significant-digit text scores (LL ~ -1xx.xxxx, gmm_ubm_kaldiHelper.py:204-208).
TOL_FRAME_LL_DELTA1 = 1e-2
TOL_SCORE_DELTA1 = 3e-4


@pytest.fixture(scope="module")
def osi(small_tree):
    from fakebob_b200.gmm_ubm_OSI import gmm_OSI
    m = gmm_OSI(small_tree["root"] + "/grp-osi", small_tree["models"], small_tree["ubm"],
                pre_model_dir=small_tree["pre_model_dir"], threshold=0.1)
    m._engine.set_debug(True)
    m._engine.set_delta_terms(3)          # the stage-by-stage tests below check the full-precision contraction
    return m


def _oracle_stages(w, cfg=None):
    from oracle import kaldi_feats as kf
    m = kf.mfcc(w)
    v = kf.compute_vad(m)
    f = kf.sliding_cmn(kf.add_deltas(m))
    return m, v, f[v != 0]


def test_frontend_stages_match_oracle(osi):
    from fakebob_b200.engine import to_audio_list
    audios = [make_audio(3, 0), make_audio(4, 1, n=24000), make_audio(5, 2, n=40001)]
    lst = to_audio_list(audios)
    osi._engine.score_avg_ll(lst)
    st = osi._engine.last_stages()
    f0 = 0
    r0 = 0
    for b, w in enumerate(lst):
        m, v, f = _oracle_stages(w)
        T = m.shape[0]
        assert st["frames"][b] == T
        gm = st["mfcc"][f0:f0 + T]
        assert np.abs(gm - m).max() < TOL_MFCC
        gv = (st["vad"][f0:f0 + T] >= 0).astype(np.float32)
        # VAD may only differ where C0 is within rounding of the threshold
        assert (gv != v).sum() <= 1
        if (gv == v).all():
            Tv = int(v.sum())
            assert st["voiced"][b] == Tv
            gf = st["feats"][r0:r0 + Tv]
            assert np.abs(gf - f).max() < TOL_FEAT
        f0 += T
        r0 += int(st["voiced"][b])


def test_frame_loglikes_match_oracle(osi, small_oracle_models):
    from fakebob_b200.engine import to_audio_list
    ubm, spk = small_oracle_models
    lst = to_audio_list([make_audio(7, 1)])
    avg = osi._engine.score_avg_ll(lst)
    st = osi._engine.last_stages()
    X = st["feats"]
    for k, g in enumerate([ubm] + spk):
        ref = g.frame_loglikes(X)
        assert np.abs(st["frame_ll"][k] - ref).max() < TOL_FRAME_LL
        assert abs(avg[0, k] - float(g.avg_loglike(X))) < TOL_AVG_LL


def test_difference_terms_against_full_precision(small_tree, small_oracle_models):
    """Shared-variance models are scored as  ll_m = ll_0 + x.(w_m - w_0) + (g_m - g_0).  With three product terms for the
    difference the result is the full-precision one; the automatic choice (here: one term, the MAP offsets are small)
    must stay within the stated deviation of it and of the oracle."""
    from fakebob_b200.engine import GmmEngine, to_audio_list
    from oracle.scorers import OracleGmmOSI
    ubm, spk = small_oracle_models
    paths = [small_tree["ubm"]] + [m[2] for m in small_tree["models"]]
    lst = to_audio_list([make_audio(8, 0), make_audio(9, 2)])
    want = OracleGmmOSI(ubm, spk).score(lst)
    res = {}
    for terms in (3, 2, 1, 0):
        eng = GmmEngine.from_files(paths, delta_terms=terms)
        info = eng.gmm_info()
        assert info["shared_variances"] and info["delta_terms"] == (terms or info["delta_terms"])
        a = eng.score_avg_ll(lst)
        res[terms] = (a, eng.last_stages()["frame_ll"].copy(), info)
        eng.close()
    a3, f3, _ = res[3]
    assert np.abs((a3[:, 1:] - a3[:, :1]) - want).max() < TOL_SCORE
    for terms in (2, 1, 0):
        a, f, info = res[terms]
        sdev = float(np.abs((a[:, 1:] - a[:, :1]) - (a3[:, 1:] - a3[:, :1])).max())
        fdev = float(np.abs(f - f3).max())
        print("difference terms %d (in effect %d, err estimate %.2e): frame LL deviation %.2e, score deviation %.2e vs three terms"
              % (terms, info["delta_terms"], info["err_estimate"], fdev, sdev))
        assert np.array_equal(f[0], f3[0])                         # slot 0 is always the three-term contraction
        # forced one / two terms: the deviation scales with the size of the MAP offsets (err_estimate predicts the per-frame
        # error of one term); the automatic choice must stay inside the stated tolerances whatever the models are
        scale = max(1.0, info["err_estimate"] / 3e-4)
        assert fdev < TOL_FRAME_LL_DELTA1 * scale and sdev < TOL_SCORE_DELTA1 * scale
        if terms == 0:
            assert fdev < TOL_FRAME_LL_DELTA1 and sdev < TOL_SCORE_DELTA1
            assert np.abs((a[:, 1:] - a[:, :1]) - want).max() < TOL_SCORE
    assert res[0][2]["delta_terms"] in (1, 2, 3) and res[0][2]["err_estimate"] > 0


def test_osi_scores_match_oracle(osi, small_oracle_models):
    from oracle.scorers import OracleGmmOSI
    ubm, spk = small_oracle_models
    ref = OracleGmmOSI(ubm, spk, threshold=0.1)
    batch = np.stack([make_audio(s, s % 3) for s in range(11, 16)], axis=1)      # (N, B) columns
    got = osi.score(batch)
    want = ref.score(batch)
    assert got.shape == want.shape == (5, 3)
    assert np.abs(got - want).max() < TOL_SCORE
    d1, s1 = osi.make_decisions(batch)
    d2, s2 = ref.make_decisions(batch)
    assert list(d1) == list(d2)
    one = osi.score(batch[:, 0])
    assert one.shape == (3,)
    assert np.abs(one - want[0]).max() < TOL_SCORE


def test_csi_sv_wrappers(small_tree, small_oracle_models):
    from fakebob_b200.gmm_ubm_CSI import gmm_CSI
    from fakebob_b200.gmm_ubm_SV import gmm_SV
    from oracle.scorers import OracleGmmCSI, OracleGmmSV
    ubm, spk = small_oracle_models
    models = small_tree["models"]
    csi = gmm_CSI(small_tree["root"] + "/grp-csi", models, pre_model_dir=small_tree["pre_model_dir"])
    ref = OracleGmmCSI(spk, [m[3] for m in models], [m[4] for m in models])
    lst = [make_audio(21, 0), make_audio(22, 1, n=20000)]
    got, want = csi.score(lst), ref.score(lst)
    assert got.shape == want.shape == (2, 3)
    assert np.abs(got - want).max() < 1e-3
    assert csi.make_decisions(lst)[0] == ref.make_decisions(lst)[0]
    sv = gmm_SV(small_tree["root"] + "/spk-sv", models[1], small_tree["ubm"], pre_model_dir=small_tree["pre_model_dir"], threshold=0.0)
    rsv = OracleGmmSV(ubm, spk[1])
    g1, w1 = sv.score(lst), rsv.score(lst)
    assert g1.shape == (2,) and np.abs(g1 - w1).max() < TOL_SCORE
    g2 = sv.score(lst[0])
    assert np.isscalar(g2) or g2.shape == ()
    dec, _ = sv.make_decisions(lst[0])
    assert dec in (1, -1)


def test_no_voiced_frames_raises(osi):
    from fakebob_b200._lib import FakebobLibraryError
    silent = np.zeros(16000, dtype=np.int16)
    silent[::7] = 1
    # constant-energy audio: every frame has the same log-energy, none exceeds 5.5 + 0.5*mean when tiny
    with pytest.raises(FakebobLibraryError):
        osi.score([silent])


def test_shared_variance_chain_matches_general_kernel(small_tree, monkeypatch):
    """MAP mean-only speaker models share the UBM's variances, so scoring uses the accumulate-chain kernel; the
    general per-model kernel (forced with FB_GMM_NO_SHARED) must give the same log-likelihoods."""
    from fakebob_b200.engine import GmmEngine, to_audio_list
    paths = [small_tree["ubm"]] + [m[2] for m in small_tree["models"]]
    lst = to_audio_list([make_audio(91, 0), make_audio(92, 1, n=20000)])
    shared = GmmEngine.from_files(paths, delta_terms=3)
    a = shared.score_avg_ll(lst)
    fa = shared.last_stages()["frame_ll"].copy()
    monkeypatch.setenv("FB_GMM_NO_SHARED", "1")
    general = GmmEngine.from_files(paths)
    b = general.score_avg_ll(lst)
    fb = general.last_stages()["frame_ll"].copy()
    assert np.abs(fa - fb).max() < 1e-3
    assert np.abs(a - b).max() < 3e-4
    # the kernels' accumulation-order offsets are common to all models and cancel in the LLR scores
    assert np.abs((a[:, 1:] - a[:, :1]) - (b[:, 1:] - b[:, :1])).max() < 5e-5
    shared.close()
    general.close()


def test_cuda_path_matches_committed_stage_fixture():
    """The CUDA front-end and GMM kernel against tests/golden/kaldi_stages.npz (a committed artefact of the oracle, see
    tests/golden/make_golden_kaldi.py) -- independent of the oracle's current code."""
    import os
    from fakebob_b200.engine import GmmEngine
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kaldi_stages.npz"))
    models = [{"weights": g["ubm_weights"], "means_invvars": g["ubm_means_invvars"], "inv_vars": g["ubm_inv_vars"], "gconsts": g["ubm_gconsts"]},
              {"weights": g["ubm_weights"], "means_invvars": g["spk_means_invvars"], "inv_vars": g["ubm_inv_vars"], "gconsts": g["spk_gconsts"]}]
    eng = GmmEngine(models, delta_terms=3)
    eng.set_debug(True)
    avg = eng.score_avg_ll([np.ascontiguousarray(g["wave"])])
    st = eng.last_stages()
    assert st["mfcc"].shape == g["mfcc"].shape and np.abs(st["mfcc"] - g["mfcc"]).max() < TOL_MFCC
    assert np.array_equal(st["vad"] >= 0, g["vad"] != 0)
    assert np.abs(st["feats"] - g["voiced_feats"]).max() < TOL_FEAT
    assert np.abs(st["frame_ll"][0] - g["frame_ll_ubm"]).max() < TOL_FRAME_LL
    assert np.abs(st["frame_ll"][1] - g["frame_ll_spk"]).max() < TOL_FRAME_LL
    assert np.abs(avg[0] - g["avg_ll"]).max() < TOL_AVG_LL
    # enrolment against the fixture's MAP-adapted model
    one = GmmEngine(models[:1])
    en = one.map_adapt([np.ascontiguousarray(g["wave"])], mean_tau=10.0)         # enrol the fixture's utterance: all voiced frames
    assert np.abs(en["occupancy"] - g["spk_all_occupancy"]).max() < 2e-2
    assert np.abs(en["means_invvars"] - g["spk_all_means_invvars"]).max() < 2e-3
    assert np.abs(en["gconsts"] - g["spk_all_gconsts"]).max() < 2e-3
    one.close()
    eng.close()


def test_evaluation_helpers_on_device_scorers_match_oracle_scorers(small_tree, small_oracle_models):
    """the numbers the reference's test.py prints (EER threshold search test.py:46-71, FRR / FAR / IER,
    closed-set accuracy) from the device scorers equal those from the oracle scorers on the same ragged audio lists."""
    from fakebob_b200 import evaluate as ev
    from fakebob_b200.gmm_ubm_CSI import gmm_CSI
    from fakebob_b200.gmm_ubm_OSI import gmm_OSI
    from fakebob_b200.gmm_ubm_SV import gmm_SV
    from oracle.scorers import OracleGmmCSI, OracleGmmOSI, OracleGmmSV
    ubm, spk = small_oracle_models
    models = small_tree["models"]
    pre = small_tree["pre_model_dir"]
    enrolled = [make_audio(9000 + k, k, n=(32000, 24000, 40000)[k]) for k in range(3)] + [make_audio(9100 + k, k) for k in range(3)]
    labels = [0, 1, 2, 0, 1, 2]
    illegal = [make_audio(700 + i, 9 + i, n=28000 + 1000 * i) for i in range(5)]
    dev = gmm_OSI(small_tree["root"] + "/ev-osi", models, small_tree["ubm"], pre_model_dir=pre)
    ref = OracleGmmOSI(ubm, spk)
    r_dev, r_ref = ev.osi_error_rates(dev, enrolled, labels, illegal), ev.osi_error_rates(ref, enrolled, labels, illegal)
    assert abs(r_dev["threshold"] - r_ref["threshold"]) < TOL_SCORE
    assert (r_dev["frr"], r_dev["ier"], r_dev["far"]) == (r_ref["frr"], r_ref["ier"], r_ref["far"])
    assert abs(dev.threshold - r_dev["threshold"]) < 1e-12            # stored on the model like test.py does
    sv_d = gmm_SV(small_tree["root"] + "/ev-sv", models[0], small_tree["ubm"], pre_model_dir=pre)
    sv_r = OracleGmmSV(ubm, spk[0])
    own = [enrolled[0], enrolled[3]]
    s_dev, s_ref = ev.sv_error_rates(sv_d, own, illegal), ev.sv_error_rates(sv_r, own, illegal)
    assert abs(s_dev["threshold"] - s_ref["threshold"]) < TOL_SCORE and (s_dev["frr"], s_dev["far"]) == (s_ref["frr"], s_ref["far"])
    csi_d = gmm_CSI(small_tree["root"] + "/ev-csi", models, pre_model_dir=pre)
    csi_r = OracleGmmCSI(spk, [m[3] for m in models], [m[4] for m in models])
    assert ev.csi_accuracy(csi_d, enrolled, labels) == ev.csi_accuracy(csi_r, enrolled, labels)
