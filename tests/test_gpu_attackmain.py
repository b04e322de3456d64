"""the reference's own driver, UNMODIFIED, on the B200 path.

baseline/_ref/attackMain.py is a byte-for-byte copy of /root/reference/attackMain.py staged by oracle/stage_reference.py
(git-ignored; /root/reference does not exist on the GPU box).  With fakebob_b200/dropin first on sys.path its imports
(attackMain.py:15-21) resolve to this package; the test fabricates the directory layout the driver hard-codes
(attackMain.py:33-36: ./model, pre-models, ./data/test-set, ./data/illegal-set), calls attackMain.main() exactly like its
__main__ block does, and checks the artefacts it writes: adversarial wavs, checkpoint pickles, and that loadData's ragged
make_decisions filter (attackMain.py:87-272) ran."""
import importlib.util
import os
import pickle
import sys

import numpy as np
import pytest

from conftest import test_audio as make_audio

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref", "attackMain.py")


@pytest.fixture(scope="module")
def driver(small_iv_tree):
    if not os.path.exists(STAGED):
        pytest.skip("baseline/_ref/attackMain.py not staged (run __graft_entry__.build() where /root/reference exists)")
    from scipy.io.wavfile import write
    t = small_iv_tree
    root = t["root"]
    spk_ids = [m[0] for m in t["models"]]
    # data/illegal-set/<imposter>/*.wav (OSI / SV), data/test-set/<enrolled spk>/*.wav (CSI); different lengths on purpose
    lengths = [32000, 24000, 40000]
    for i, name in enumerate(["imp-a", "imp-b"]):
        d = os.path.join(root, "data", "illegal-set", name)
        os.makedirs(d, exist_ok=True)
        write(os.path.join(d, "utt%d.wav" % i), 16000, (make_audio(500 + i, 5 + i, n=lengths[i]) * 32768).astype(np.int16))
    for k, s in enumerate(spk_ids):
        d = os.path.join(root, "data", "test-set", s)
        os.makedirs(d, exist_ok=True)
        write(os.path.join(d, "test%d.wav" % k), 16000, (make_audio(9000 + k, k, n=lengths[k % 3]) * 32768).astype(np.int16))
    dropin = os.path.join(ROOT, "fakebob_b200", "dropin")
    old_path, old_cwd = list(sys.path), os.getcwd()
    sys.path.insert(0, dropin)
    for mod in ("FAKEBOB", "gmm_ubm_CSI", "gmm_ubm_OSI", "gmm_ubm_SV", "ivector_PLDA_CSI", "ivector_PLDA_OSI", "ivector_PLDA_SV"):
        sys.modules.pop(mod, None)
    os.chdir(root)
    spec = importlib.util.spec_from_file_location("attackMain", STAGED)
    am = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(am)
    import FAKEBOB as dropin_fakebob
    assert dropin_fakebob.FakeBob.__module__.startswith("fakebob_b200"), "the driver must be running this package's FakeBob"
    yield {"am": am, "root": root, "spk_ids": spk_ids, "tree": t}
    os.chdir(old_cwd)
    sys.path[:] = old_path


def _main(am, spk_ids, archi, task, attack_type, threshold, max_iter=6, samples=6):
    # the argument list of attackMain.main (attackMain.py:274-276), values as its __main__ block passes them
    am.main(spk_ids, archi, task, threshold, attack_type, 0.0, 0.002, max_iter, 0.001, 1e-6, samples, 0.001, 0.9, 5, 2.0, 1, False)


def _check_outputs(root, run_id, expect_min):
    adv_dir = os.path.join(root, "adversarial-audio", run_id)
    cp_dir = os.path.join(root, "checkpoint", run_id)
    from scipy.io.wavfile import read
    wavs = [os.path.join(dp, f) for dp, _, fs in os.walk(adv_dir) for f in fs if f.endswith(".wav")]
    cps = [os.path.join(dp, f) for dp, _, fs in os.walk(cp_dir) for f in fs if f.endswith(".cp")]
    assert len(wavs) >= expect_min and len(cps) == len(wavs), (wavs, cps)
    for w in wavs:
        fs_, a = read(w)
        assert fs_ == 16000 and a.dtype == np.int16 and a.ndim in (1, 2) and a.shape[0] >= 8000
    for c in cps:
        with open(c, "rb") as f:
            rows = pickle.load(f)
        assert len(rows) >= 1 and len(rows[0]) == 4            # [distance, adver_loss, score, used_time] (FAKEBOB.py:209-217)
    return wavs


@pytest.mark.timeout(300)
def test_reference_driver_gmm_osi_targeted(driver, capsys):
    """archi gmm, task OSI, targeted: loadData keeps the imposter audios the system rejects (decision == -1), estimate_threshold
    runs on one of them, then one attack per (audio, target speaker)."""
    from fakebob_b200.gmm_ubm_OSI import gmm_OSI
    am, t = driver["am"], driver["tree"]
    probe = gmm_OSI(t["root"] + "/probe-osi", t["models"], t["ubm"], pre_model_dir=t["pre_model_dir"])
    from scipy.io.wavfile import read
    files = sorted(os.path.join(dp, f) for dp, _, fs in os.walk(os.path.join(t["root"], "data", "illegal-set")) for f in fs)
    scores = probe.score([read(f)[1] for f in files])              # ragged list of int16 audios
    thr = float(np.max(scores)) + 0.03                             # every imposter is rejected, and reachable by the attack
    _main(am, driver["spk_ids"], "gmm", "OSI", "targeted", thr, max_iter=6)
    out = capsys.readouterr().out
    assert "------ load data done, total num: %d ------" % (len(files) * len(driver["spk_ids"])) in out
    assert "return at iter_outer" in out                           # estimate_threshold finished (FAKEBOB.py:97-100)
    wavs = _check_outputs(t["root"], "gmm-OSI-targeted", len(files) * len(driver["spk_ids"]))
    assert any(w.endswith("_2.wav") for w in wavs)                 # <name>_<target label>.wav (attackMain.py:238)


@pytest.mark.timeout(300)
def test_reference_driver_gmm_csi_untargeted_filters_ragged_list(driver, capsys):
    """task CSI: loadData scores the whole test set as ONE ragged list and keeps the correctly classified audios."""
    from fakebob_b200.gmm_ubm_CSI import gmm_CSI
    am, t = driver["am"], driver["tree"]
    probe = gmm_CSI(t["root"] + "/probe-csi", t["models"], pre_model_dir=t["pre_model_dir"])
    from scipy.io.wavfile import read
    kept = 0
    for k, s in enumerate(probe.spk_ids):
        d = os.path.join(t["root"], "data", "test-set", s)
        for f in os.listdir(d):
            dec, _ = probe.make_decisions(read(os.path.join(d, f))[1])
            kept += int(dec == k)
    _main(am, driver["spk_ids"], "gmm", "CSI", "untargeted", 0.0, max_iter=4)
    out = capsys.readouterr().out
    assert "------ load data done, total num: %d ------" % kept in out
    if kept:
        _check_outputs(t["root"], "gmm-CSI-untargeted", kept)


@pytest.mark.timeout(300)
def test_reference_driver_iv_sv(driver, capsys):
    """archi iv, task SV: iv_SV(id, model, threshold=...) with the default pre_model_dir, estimate_threshold, attack."""
    from fakebob_b200.ivector_PLDA_SV import iv_SV
    am, t = driver["am"], driver["tree"]
    spk = sorted(t["iv_models"], key=lambda m: m[0])[0]
    probe = iv_SV(t["root"] + "/probe-ivsv", spk, pre_model_dir=t["pre_model_dir"])
    from scipy.io.wavfile import read
    files = sorted(os.path.join(dp, f) for dp, _, fs in os.walk(os.path.join(t["root"], "data", "illegal-set")) for f in fs)
    scores = np.atleast_1d(probe.score([read(f)[1] for f in files]))
    thr = float(np.max(scores)) + 0.05 * max(1.0, abs(float(np.max(scores))))
    _main(am, [spk[0]], "iv", "SV", "targeted", thr, max_iter=4)
    out = capsys.readouterr().out
    assert "------ load data done, total num: %d ------" % len(files) in out
    _check_outputs(t["root"], os.path.join("iv-SV-targeted", spk[0]), len(files))
