"""The NES-loop oracle (oracle/nes.py) is PINNED here: against golden trajectories produced by the reference's
own FAKEBOB.py (tests/golden/make_golden.py) and, when /root/reference is present, against a live run of it."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from stub_scorer import CASES, StubScorer, make_audio  # noqa: E402

from oracle.nes import OracleFakeBob, margin_loss  # noqa: E402


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(name):
    task, attack_type, K, kw, hp = CASES[name]
    gold = np.load(os.path.join(HERE, "golden", "nes_%s.npz" % name))
    model = StubScorer(K, 4000, sv=(task == "SV"))
    np.random.seed(2024)
    fb = OracleFakeBob(task, attack_type, model, **hp)
    adver, flag = fb.attack(make_audio(1).copy(), None, **kw)
    assert flag == int(gold["flag"])
    assert len(fb.log) == int(gold["n_rows"])
    assert np.array_equal(adver, gold["adver"])                     # bit-exact int16 adversarial audio
    assert np.array_equal(np.array([r[0] for r in fb.log]), gold["distance"])
    assert np.array_equal(np.array([float(np.asarray(r[1]).reshape(-1)[0]) for r in fb.log]), gold["adver_loss"])


def test_oracle_estimate_threshold_matches_reference_golden():
    gold = np.load(os.path.join(HERE, "golden", "nes_estimate_threshold.npz"))
    model = StubScorer(3, 4000, threshold=2.45)
    np.random.seed(7)
    fb = OracleFakeBob("OSI", "targeted", model, max_iter=10, samples_per_draw=10)
    res = fb.estimate_threshold(make_audio(2))
    assert res[0] == float(gold["score"]) and res[1] == int(gold["n_iters"])
    assert fb.attack_type == "targeted"


@pytest.mark.skipif(not os.path.exists("/root/reference/FAKEBOB.py"), reason="reference checkout not present")
def test_oracle_matches_live_reference():
    sys.path.insert(0, "/root/reference")
    import FAKEBOB as REF
    task, attack_type, K, kw, hp = CASES["osi_untargeted"]
    hp = dict(hp, samples_per_draw=6, max_iter=15)
    model = StubScorer(K, 4000)
    np.random.seed(11)
    ref = REF.FakeBob(task, attack_type, model, **hp)
    import tempfile
    with contextlib.redirect_stdout(io.StringIO()):
        a1, f1 = ref.attack(make_audio(3).copy(), os.path.join(tempfile.mkdtemp(), "cp"), **kw)
        np.random.seed(12)
        g_ref = ref.get_grad(make_audio(4))
    np.random.seed(11)
    ora = OracleFakeBob(task, attack_type, model, **hp)
    a2, f2 = ora.attack(make_audio(3).copy(), None, **kw)
    ora.threshold = ref.threshold
    np.random.seed(12)
    g_ora = ora.get_grad(make_audio(4))
    assert f1 == f2 and np.array_equal(a1, a2)
    assert g_ref[0] == g_ora[0] and np.array_equal(g_ref[1], g_ora[1])


def test_margin_loss_variants():
    s = np.array([[0.1, 0.5, -0.2], [0.3, 0.0, 0.9]])
    assert np.allclose(margin_loss(s, "OSI", "untargeted", 1.0, 0.1), [[0.6], [0.2]])
    assert np.allclose(margin_loss(s, "OSI", "targeted", 0.2, 0.0, target=0), [[0.4], [0.6]])
    assert np.allclose(margin_loss(s, "CSI", "targeted", 0.0, 0.0, target=2), [[0.7], [-0.6]])
    assert np.allclose(margin_loss(s, "CSI", "untargeted", 0.0, 0.0, true=1), [[0.4], [-0.9]])
    assert np.allclose(margin_loss(np.array([0.4, 0.7]), "SV", "untargeted", 0.5, 0.0), [[0.1], [-0.2]])
