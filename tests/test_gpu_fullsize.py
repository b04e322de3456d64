"""Full-size (BASELINE.json configs[1]: 51 audios x 5 s, UBM + 5 speaker GMMs x 2048 mixtures, samples_per_draw = 50)
checks through size-independent properties -- the CPU oracle needs minutes at this size, so it is not run here:
batch independence, permutation invariance, the tensor-core kernel against the fp32 CUDA-core cross-check kernel,
determinism of the Philox NES loop, and the L-infinity box of the update."""
import os

import numpy as np
import pytest

from conftest import test_audio as make_audio

pytestmark = pytest.mark.gpu

C, K, S, N = 2048, 5, 50, 80000


def _models(seed=0):
    """UBM-like GMM + K 'MAP-adapted' copies (shared variances, shifted means), scaled to the feature statistics."""
    r = np.random.default_rng(seed)
    iv = r.uniform(0.05, 0.6, (C, 72)).astype(np.float32)
    mu = (r.standard_normal((C, 72)) * 2.5).astype(np.float32)
    w = r.dirichlet(np.full(C, 5.0)).astype(np.float32)

    def gmm(m):
        gc = (np.log(w.astype(np.float64)) - 0.5 * (72 * np.log(2 * np.pi) - np.log(iv.astype(np.float64)).sum(1)
                                                     + (m.astype(np.float64) ** 2 * iv).sum(1))).astype(np.float32)
        return {"weights": w, "means_invvars": (m * iv).astype(np.float32), "inv_vars": iv, "gconsts": gc}
    return [gmm(mu)] + [gmm(mu + 0.05 * r.standard_normal((C, 72)).astype(np.float32)) for _ in range(K)]


@pytest.fixture(scope="module")
def engine():
    from fakebob_b200.engine import GmmEngine
    e = GmmEngine(_models())
    yield e
    e.close()


@pytest.fixture(scope="module")
def batch():
    from fakebob_b200.engine import to_audio_list
    return to_audio_list([make_audio(300 + i, i % 7, n=N) for i in range(S + 1)])


def test_batch_independence_and_permutation(engine, batch):
    full = engine.score_avg_ll(batch)
    assert full.shape == (S + 1, K + 1) and np.isfinite(full).all()
    rows = engine.voiced_rows()
    assert 0.5 * (S + 1) * 500 < rows <= (S + 1) * 500
    for i in (0, 17, S):
        alone = engine.score_avg_ll([batch[i]])
        assert np.abs(alone[0] - full[i]).max() < 2e-4          # same frames, different tiles / segment cuts
    rev = engine.score_avg_ll(batch[::-1])
    assert np.abs(rev[::-1] - full).max() < 2e-4
    # ragged: different lengths in one batch (loadData passes such lists, attackMain.py:128)
    ragged = [batch[0][:31234], batch[1], batch[2][:16000]]
    r = engine.score_avg_ll(ragged)
    assert np.abs(r[1] - full[1]).max() < 2e-4
    assert np.abs(engine.score_avg_ll([ragged[0]])[0] - r[0]).max() < 2e-4


def test_tensor_core_kernel_matches_fp32_cross_check(engine, batch):
    a = engine.score_avg_ll(batch)
    fa = engine.last_stages()["frame_ll"].copy()
    engine.set_gmm_impl("simt")
    try:
        b = engine.score_avg_ll(batch)
        fb = engine.last_stages()["frame_ll"].copy()
    finally:
        engine.set_gmm_impl("umma")
    assert fa.shape == fb.shape and fa.shape[0] == K + 1
    # per-frame LL: relative 5e-6 (the random models sit far from the features, |LL| is 500-900 here; the small-tree
    # tests hold 2e-3 absolute on |LL| ~ 120)
    assert (np.abs(fa - fb) / np.abs(fb)).max() < 5e-6
    assert (np.abs(a - b) / np.abs(b)).max() < 2e-6
    assert np.abs((a[:, 1:] - a[:, :1]) - (b[:, 1:] - b[:, :1])).max() < 3e-4


def test_nes_full_size_deterministic_and_boxed(tmp_path):
    from fakebob_b200 import kaldi_io
    from fakebob_b200.FAKEBOB import FakeBob
    from fakebob_b200.gmm_ubm_OSI import gmm_OSI
    from fakebob_b200.synth import write_conf
    pre = str(tmp_path / "pre-models")
    os.makedirs(pre)
    write_conf(pre)
    ms = _models(1)
    paths = []
    for i, g in enumerate(ms):
        p = os.path.join(pre, "final.dubm" if i == 0 else "spk%d-identity.gmm" % i)
        kaldi_io.write_diag_gmm(p, g["weights"], g["means_invvars"], g["inv_vars"], g["gconsts"])
        paths.append(os.path.abspath(p))
    models = [["%04d" % (1000 + i), "u%d" % i, paths[i], 0.0, 1.0] for i in range(1, K + 1)]
    model = gmm_OSI(str(tmp_path / "grp"), models, paths[0], pre_model_dir=pre)
    audio = make_audio(400, 3, n=N)
    eps = 0.002
    outs = []
    for _ in range(2):
        fb = FakeBob("OSI", "untargeted", model, epsilon=eps, max_iter=4, samples_per_draw=S, seed=99, verbose=False)
        adv, flag = fb.attack(audio.copy(), None, threshold=1e3)
        outs.append((adv.copy(), flag, fb.log.copy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1] == -1
    assert np.array_equal(outs[0][2], outs[1][2])                # distance / loss / lr / scores per iteration
    adv = outs[0][0]
    assert adv.dtype == np.int16 and adv.shape == (N, 1)
    # L-infinity box of FAKEBOB.py:202-203 in int16 units (truncation adds < 1)
    assert np.abs(adv[:, 0].astype(np.int64) - (audio * 32768).astype(np.int64)).max() <= int(eps * 32768) + 1
    log = outs[0][2]
    assert log.shape[0] == 4 and np.all(log[:, 0] <= eps + 1e-12)
