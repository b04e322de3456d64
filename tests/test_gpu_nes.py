"""GPU parity tests for the NES attack loop (FakeBob.attack / get_grad / estimate_threshold)."""
import os
import pickle

import numpy as np
import pytest

from conftest import test_audio as make_audio

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scorers(small_tree):
    from fakebob_b200.gmm_ubm_CSI import gmm_CSI
    from fakebob_b200.gmm_ubm_OSI import gmm_OSI
    from fakebob_b200.gmm_ubm_SV import gmm_SV
    t = small_tree
    return {
        "OSI": gmm_OSI(t["root"] + "/n-osi", t["models"], t["ubm"], pre_model_dir=t["pre_model_dir"]),
        "CSI": gmm_CSI(t["root"] + "/n-csi", t["models"], pre_model_dir=t["pre_model_dir"]),
        "SV": gmm_SV(t["root"] + "/n-sv", t["models"][0], t["ubm"], pre_model_dir=t["pre_model_dir"]),
    }


CASES = [
    ("OSI", "untargeted", dict(threshold=1.0)),
    ("OSI", "targeted", dict(threshold=0.05, target=2)),
    ("CSI", "targeted", dict(target=1)),
    ("CSI", "untargeted", dict(true=0)),
    ("SV", "untargeted", dict(threshold=1.0)),
]


@pytest.mark.parametrize("task,attack_type,kw", CASES)
def test_attack_bit_exact_given_same_scores_numpy_rng(scorers, tmp_path, task, attack_type, kw):
    """Layer (ii) of SURVEY hard-part 3: with the reference's numpy noise stream and the same scorer, the
    device loop (perturb, int16 quantisation, loss, pairwise-order gradient, momentum, sign step, clip,
    plateau schedule, early stop) is bit-identical to the CPU restatement of FAKEBOB.py driven by that scorer."""
    from fakebob_b200.FAKEBOB import FakeBob
    from oracle.nes import OracleFakeBob
    model = scorers[task]
    audio = make_audio(31, 0, n=16000)
    hp = dict(adver_thresh=0.0, epsilon=0.002, max_iter=12, samples_per_draw=10, plateau_length=3)
    np.random.seed(99)
    fb = FakeBob(task, attack_type, model, rng="numpy", verbose=False, **hp)
    cp = str(tmp_path / "cp.pkl")
    adv_g, flag_g = fb.attack(audio.copy(), cp, **kw)
    np.random.seed(99)                       # rng="numpy": FakeBob draws nothing at construction, same stream as the reference
    ob = OracleFakeBob(task, attack_type, model, **hp)
    adv_o, flag_o = ob.attack(audio.copy(), None, **kw)
    assert flag_g == flag_o
    assert len(ob.log) == fb.iters_done
    assert np.array_equal(fb.final_adver, ob.final_adver)
    assert np.array_equal(adv_g, adv_o)
    for it, row in enumerate(ob.log):
        assert fb.log[it, 0] == row[0]                       # distance
        assert fb.log[it, 1] == float(np.asarray(row[1]).reshape(-1)[0])   # adver_loss
    with open(cp, "rb") as f:
        rows = pickle.load(f)
    assert len(rows) == fb.iters_done and len(rows[0]) == 4


def test_attack_philox_matches_oracle_philox(scorers):
    """rng='philox': the device generator and oracle/philox.py produce the same stream (integers exact,
    Box-Muller within libm ulps), so with the same scorer the trajectories coincide."""
    from fakebob_b200.FAKEBOB import FakeBob
    from oracle.nes import OracleFakeBob, PhiloxNoise
    model = scorers["OSI"]
    audio = make_audio(32, 1, n=16000)
    hp = dict(max_iter=10, samples_per_draw=8, plateau_length=3)
    fb = FakeBob("OSI", "untargeted", model, rng="philox", seed=0x1234ABCD5678, verbose=False, iters_per_launch=4, **hp)
    adv_g, flag_g = fb.attack(audio.copy(), None, threshold=1.0)
    ob = OracleFakeBob("OSI", "untargeted", model, noise_fn=PhiloxNoise(0x1234ABCD5678), **hp)
    adv_o, flag_o = ob.attack(audio.copy(), None, threshold=1.0)
    assert flag_g == flag_o and fb.iters_done == len(ob.log)
    agree = np.mean(fb.final_adver == ob.final_adver)
    assert agree > 0.9999
    assert np.abs(fb.log[:len(ob.log), 1] - np.array([float(np.asarray(r[1]).reshape(-1)[0]) for r in ob.log])).max() < 1e-6


def test_early_stop_and_success_flag(scorers):
    from fakebob_b200.FAKEBOB import FakeBob
    model = scorers["SV"]
    audio = make_audio(33, 0, n=16000)
    s0 = model.score(audio)
    fb = FakeBob("SV", "untargeted", model, max_iter=20, samples_per_draw=6, verbose=False, seed=1)
    adv, flag = fb.attack(audio, None, threshold=float(s0) - 1.0)      # already accepted -> loss < 0 at iter 0
    assert flag == 1 and fb.iters_done == 1
    assert adv.shape == (16000, 1) and adv.dtype == np.int16
    assert np.array_equal(adv[:, 0], (audio * 32768).astype(np.int16))
    fb2 = FakeBob("SV", "untargeted", model, max_iter=5, samples_per_draw=6, verbose=False, seed=1)
    adv2, flag2 = fb2.attack(audio, None, threshold=float(s0) + 50.0)  # unreachable
    assert flag2 == -1 and fb2.iters_done == 5
    assert np.max(np.abs(fb2.final_adver[:, 0] - audio)) <= 0.002 + 1e-12


def test_end_to_end_against_cpu_oracle_scorer(scorers, small_oracle_models):
    """Layer (iii): GPU attack vs the all-CPU oracle (oracle scorer + oracle NES), same Philox stream.
    Scores differ at the 1e-4 level so trajectories are compared statistically."""
    from fakebob_b200.FAKEBOB import FakeBob
    from oracle.nes import OracleFakeBob, PhiloxNoise
    from oracle.scorers import OracleGmmSV
    ubm, spk = small_oracle_models
    model = scorers["SV"]
    ref_model = OracleGmmSV(ubm, spk[0])
    audio = make_audio(34, 2, n=16000)
    thr = float(model.score(audio)) + 0.08
    hp = dict(max_iter=8, samples_per_draw=8)
    fb = FakeBob("SV", "untargeted", model, seed=77, verbose=False, **hp)
    fb.attack(audio.copy(), None, threshold=thr)
    ob = OracleFakeBob("SV", "untargeted", ref_model, noise_fn=PhiloxNoise(77), **hp)
    ob.attack(audio.copy(), None, threshold=thr)
    n = min(fb.iters_done, len(ob.log))
    assert abs(fb.iters_done - len(ob.log)) <= 1
    lg = fb.log[:n, 1]
    lo = np.array([float(np.asarray(r[1]).reshape(-1)[0]) for r in ob.log[:n]])
    assert np.abs(lg - lo).max() < 5e-3
    assert np.mean(np.sign(fb.final_adver - audio[:, None]) == np.sign(ob.final_adver - audio[:, None])) > 0.8


def test_get_grad_and_estimate_threshold(scorers):
    from fakebob_b200.FAKEBOB import FakeBob
    from oracle.nes import OracleFakeBob
    model = scorers["OSI"]
    audio = make_audio(35, 1, n=16000)
    np.random.seed(5)
    fb = FakeBob("OSI", "untargeted", model, samples_per_draw=8, rng="numpy", verbose=False)
    fb.threshold = 0.7
    fl, g, al, sc = fb.get_grad(audio)
    np.random.seed(5)
    ob = OracleFakeBob("OSI", "untargeted", model, samples_per_draw=8)
    ob.threshold = 0.7
    fl2, g2, al2, sc2 = ob.get_grad(audio)
    assert fl == fl2 and np.array_equal(g, g2) and al[0] == al2[0] and np.array_equal(sc, sc2)
    # estimate_threshold: same control flow and result as the restated reference with the same scorer
    model.threshold = float(np.max(model.score(audio))) + 0.02
    np.random.seed(6)
    fb = FakeBob("OSI", "targeted", model, samples_per_draw=8, rng="numpy", verbose=False, max_lr=0.001)
    r1 = fb.estimate_threshold(audio)
    np.random.seed(6)
    ob = OracleFakeBob("OSI", "targeted", model, samples_per_draw=8, max_lr=0.001)
    r2 = ob.estimate_threshold(audio)
    assert r1[0] == r2[0] and r1[1] == r2[1]
    assert fb.attack_type == "targeted"
    model.threshold = 0.0


def test_back_to_back_sessions_with_different_length_and_draw(scorers):
    """Consecutive attacks with no score() call in between (attackMain.py's loops, sharding.attack_many) on audios of
    different length and with different samples_per_draw: every session must lay the shared batch workspace out afresh."""
    from fakebob_b200.FAKEBOB import FakeBob
    from oracle.nes import OracleFakeBob, PhiloxNoise
    model = scorers["OSI"]
    plan = [(make_audio(36, 0, n=16000), 6, 5), (make_audio(37, 1, n=24000), 10, 6), (make_audio(38, 2, n=8000), 4, 7),
            (make_audio(39, 0, n=24000), 10, 8)]
    got = []
    for audio, S, seed in plan:                          # device sessions back to back, nothing else touches the context
        fb = FakeBob("OSI", "untargeted", model, max_iter=3, samples_per_draw=S, seed=seed, verbose=False)
        adv, flag = fb.attack(audio.copy(), None, threshold=1.0)
        got.append((adv.copy(), flag, fb.log.copy()))
    for (audio, S, seed), (adv, flag, log) in zip(plan, got):
        ob = OracleFakeBob("OSI", "untargeted", model, max_iter=3, samples_per_draw=S, noise_fn=PhiloxNoise(seed))
        adv_o, flag_o = ob.attack(audio.copy(), None, threshold=1.0)
        assert flag == flag_o and log.shape[0] == len(ob.log)
        assert adv.shape == adv_o.shape and np.mean(adv == adv_o) > 0.9999
        assert np.abs(log[:, 1] - np.array([float(np.asarray(r[1]).reshape(-1)[0]) for r in ob.log])).max() < 1e-6


def test_device_error_is_reported_once_with_its_own_code(scorers):
    from fakebob_b200._lib import FakebobLibraryError
    from fakebob_b200.FAKEBOB import FakeBob
    model = scorers["SV"]
    silent = np.zeros(16000)
    silent[::7] = 1.0 / 32768
    fb = FakeBob("SV", "untargeted", model, max_iter=2, samples_per_draw=4, seed=1, verbose=False, sigma=1e-9)
    with pytest.raises(FakebobLibraryError) as ei:
        fb.attack(silent, None, threshold=1e3)
    assert "error -4" in str(ei.value) and "no voiced frames" in str(ei.value)
    # the flag was cleared: an unrelated batch right after succeeds
    ok = model.score(make_audio(33, 0, n=16000))
    assert np.isfinite(ok)


@pytest.mark.parametrize("task", ["OSI", "SV"])
def test_estimate_threshold_on_device_matches_oracle_search(scorers, task, capsys):
    """rng='philox': the whole inner loop of FAKEBOB.py:76-137 runs on the device (the clean column of each NES batch is the
    make_decisions() score); same Philox stream and scorer as the restated reference -> same number of iterations, same
    returned score up to the batch-composition tolerance of the scorer."""
    from fakebob_b200.FAKEBOB import FakeBob
    from oracle.nes import OracleFakeBob, PhiloxNoise
    model = scorers[task]
    audio = make_audio(41, 1, n=16000)
    s0 = model.score(audio)
    model.threshold = float(np.max(s0)) + 0.02                      # rejected now, reachable within the epsilon ball
    try:
        fb = FakeBob(task, "targeted", model, samples_per_draw=8, seed=4242, verbose=True, max_lr=0.001, iters_per_launch=4)
        r1 = fb.estimate_threshold(audio)
        out = capsys.readouterr().out
        ob = OracleFakeBob(task, "targeted", model, samples_per_draw=8, max_lr=0.001, noise_fn=PhiloxNoise(4242))
        r2 = ob.estimate_threshold(audio)
    finally:
        model.threshold = 0.0
    assert fb.attack_type == "targeted"
    assert "----- iter_outer:0" in out and "return at iter_outer" in out and "cost %d iters" % r1[1] in out
    assert r1[1] == r2[1] and r1[1] >= 1                           # inner iterations with an update
    assert abs(float(r1[0]) - float(r2[0])) < 5e-4
    assert abs(fb.threshold - ob.threshold) < 5e-4 and fb.draws == ob.draws


class _BlackBox(object):
    """A foreign system in the sense of the reference's README.md:136: only score() / make_decisions(), nothing else."""

    def __init__(self, inner):
        self._inner = inner
        self.threshold = getattr(inner, "threshold", 0.0)
        self.calls = 0

    def score(self, audios, fs=16000, bits_per_sample=16, n_jobs=5, debug=False):
        self.calls += 1
        return self._inner.score(audios, fs=fs, bits_per_sample=bits_per_sample)

    def make_decisions(self, audios, fs=16000, bits_per_sample=16, n_jobs=5, debug=False):
        self._inner.threshold = self.threshold
        return self._inner.make_decisions(audios, fs=fs, bits_per_sample=bits_per_sample)


@pytest.mark.parametrize("task,attack_type,kw", [("OSI", "targeted", dict(threshold=0.05, target=2)), ("SV", "untargeted", dict(threshold=1.0)),
                                                 ("CSI", "untargeted", dict(true=0))])
def test_black_box_model_attack_equals_resident_model_attack(scorers, task, attack_type, kw):
    """FakeBob on a duck-typed scorer: NES state, noise, quantisation, loss, gradient and update on the device, scoring by
    the caller's score().  Wrapping one of this package's scorers as a black box must reproduce the all-device attack."""
    from fakebob_b200.FAKEBOB import FakeBob
    inner = scorers[task]
    box = _BlackBox(inner)
    assert not hasattr(box, "_engine")
    audio = make_audio(45, 1, n=16000)
    hp = dict(max_iter=5, samples_per_draw=8, plateau_length=3)
    fa = FakeBob(task, attack_type, inner, seed=31, verbose=False, **hp)
    adv_a, flag_a = fa.attack(audio.copy(), None, **kw)
    fb = FakeBob(task, attack_type, box, seed=31, verbose=False, **hp)
    adv_b, flag_b = fb.attack(audio.copy(), None, **kw)
    assert flag_a == flag_b and fa.iters_done == fb.iters_done and box.calls >= fb.iters_done
    assert np.array_equal(adv_a, adv_b)
    assert np.allclose(fa.log[:, :3], fb.log[:, :3], rtol=0, atol=1e-12)
    # get_grad through the black box
    fl, g, al, sc = fb.get_grad(audio)
    assert g.shape == (16000, 1) and np.isfinite(g).all() and np.isfinite(fl)


def test_black_box_estimate_threshold(scorers):
    from fakebob_b200.FAKEBOB import FakeBob
    inner = scorers["SV"]
    audio = make_audio(46, 2, n=16000)
    box = _BlackBox(inner)
    box.threshold = float(inner.score(audio)) + 0.02
    try:
        fb = FakeBob("SV", "targeted", box, samples_per_draw=8, seed=77, verbose=False, max_lr=0.001)
        score, n_iters, secs = fb.estimate_threshold(audio)
    finally:
        inner.threshold = 0.0
    assert score >= box.threshold and n_iters >= 1 and fb.attack_type == "targeted"


def test_model_must_have_the_scorer_interface():
    from fakebob_b200.FAKEBOB import FakeBob

    class Stub:
        def score(self, a):
            return 0.0

    with pytest.raises(TypeError):
        FakeBob("SV", "untargeted", Stub())                         # no make_decisions()
