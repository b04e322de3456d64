"""the non-ideal effects of the reference's real Kaldi path on the device (fb_set_kaldi_exact) against
the oracle switches (oracle/kaldi_nonideal.py): CompressedMatrix round trip of the MFCCs, 7-significant-digit text values."""
import numpy as np
import pytest

from conftest import test_audio as make_audio

pytestmark = pytest.mark.gpu


def _sig7_equal(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= 2e-7 * np.maximum(np.abs(a), np.abs(b)) + 1e-300)


def test_compressed_features_and_text_scores_gmm(small_tree, small_oracle_models):
    from fakebob_b200.engine import GmmEngine, to_audio_list
    from oracle import kaldi_feats as kf
    from oracle import kaldi_nonideal as kn
    ubm, spk = small_oracle_models
    gmms = [ubm] + spk
    paths = [small_tree["ubm"]] + [m[2] for m in small_tree["models"]]
    lst = to_audio_list([make_audio(51, 0), make_audio(52, 1, n=24000)])
    eng = GmmEngine.from_files(paths, delta_terms=3)
    eng.set_debug(True)
    plain = eng.score_avg_ll(lst)
    st0 = eng.last_stages()
    eng.set_kaldi_exact(compress=True, text=True)
    got = eng.score_avg_ll(lst)
    st1 = eng.last_stages()
    assert _sig7_equal(got, kn.round_sig7(got))                     # what the reference would have parsed from text
    f0 = 0
    for u in range(len(lst)):
        T = int(st0["frames"][u])
        m_plain = st0["mfcc"][f0:f0 + T]
        want_m = kn.compress_decompress(m_plain)                    # the oracle codec on the device's own MFCCs
        got_m = st1["mfcc"][f0:f0 + T]
        step = float(m_plain.max() - m_plain.min()) / 63.0
        close = np.abs(got_m - want_m) < 1e-4
        assert close.mean() > 0.999 and np.abs(got_m - want_m).max() < step    # a byte may flip at an exact rounding boundary
        assert 0.01 * step < np.abs(got_m - m_plain).max() <= 0.51 * step + 1e-3     # the codec did something, and only that
        # downstream of the compressed matrix the oracle pipeline must give the device's (rounded) average log-likelihoods
        v = kf.compute_vad(got_m)
        X = kf.select_voiced(kf.sliding_cmn(kf.add_deltas(got_m)), v)
        want = np.array([kn.round_sig7(float(g.avg_loglike(X))) for g in gmms])
        assert np.abs(got[u] - want).max() < 5e-4
        f0 += T
    assert np.abs(got - plain).max() > 1e-3                          # compression moves the scores far more than any rounding
    eng.set_kaldi_exact(False, False)
    assert np.array_equal(eng.score_avg_ll(lst), plain)
    eng.close()


def test_text_precision_ivector_path(small_iv_tree):
    from fakebob_b200.engine import to_audio_list
    from fakebob_b200.ivector_PLDA_OSI import iv_OSI
    from oracle import kaldi_nonideal as kn
    t = small_iv_tree
    m = iv_OSI(t["root"] + "/iv-osi-kx", t["iv_models"], pre_model_dir=t["pre_model_dir"])
    lst = to_audio_list([make_audio(62, 0), make_audio(63, 1, n=24000)])
    plain, iv_plain = m._engine.score_plda(lst, want_ivectors=True)
    m._engine.set_kaldi_exact(compress=False, text=True)
    got, iv = m._engine.score_plda(lst, want_ivectors=True)
    assert _sig7_equal(got, kn.round_sig7(got)) and _sig7_equal(iv, kn.round_sig7_f32(iv))
    assert np.abs(iv - kn.round_sig7_f32(iv_plain)).max() <= 2e-7 * np.abs(iv_plain).max() + 1e-12
    order = np.argsort([x[0] for x in t["iv_models"]])
    want = kn.round_sig7(t["system"].plda_scores(t["enrolled"][order], kn.round_sig7_f32(iv_plain)))
    assert np.abs(got - want).max() < 2e-3
    assert np.abs(got - plain).max() < 1e-2
    m._engine.set_kaldi_exact(False, False)


def test_nes_loop_runs_in_kaldi_exact_mode(small_tree):
    """The attack loop with both switches on: scores inside the captured iteration are the rounded ones."""
    from fakebob_b200.FAKEBOB import FakeBob
    from fakebob_b200.gmm_ubm_SV import gmm_SV
    from oracle.nes import OracleFakeBob, PhiloxNoise
    sv = gmm_SV(small_tree["root"] + "/sv-kx", small_tree["models"][0], small_tree["ubm"], pre_model_dir=small_tree["pre_model_dir"])
    sv._engine.set_kaldi_exact(compress=True, text=True)
    audio = make_audio(53, 2, n=16000)
    s0 = float(sv.score(audio))
    hp = dict(max_iter=4, samples_per_draw=6)
    fb = FakeBob("SV", "untargeted", sv, seed=9, verbose=False, **hp)
    adv, flag = fb.attack(audio, None, threshold=s0 + 5.0)
    ob = OracleFakeBob("SV", "untargeted", sv, noise_fn=PhiloxNoise(9), **hp)     # same (device) scorer, oracle loop
    adv_o, flag_o = ob.attack(audio, None, threshold=s0 + 5.0)
    assert flag == flag_o == -1 and fb.iters_done == 4
    assert np.mean(adv == adv_o) > 0.999
    assert abs(fb.log[0, 4] - s0) < 1e-9
