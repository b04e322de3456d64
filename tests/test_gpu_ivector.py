"""GPU parity tests for the i-vector / PLDA scoring path (fb_score_ivector_host and the iv_* wrappers)."""
import numpy as np
import pytest

from conftest import test_audio as make_audio

pytestmark = pytest.mark.gpu

TOL_POST = 2e-3        # absolute, pruned/renormalised posteriors of components both sides keep
TOL_IVEC = 5e-3        # absolute, raw i-vector elements (|w| up to ~20 incl. the prior-offset dimension)
TOL_PLDA = 2e-2        # absolute, PLDA log-likelihood ratio (|LLR| ~ 20)


@pytest.fixture(scope="module")
def iv_osi(small_iv_tree):
    from fakebob_b200.ivector_PLDA_OSI import iv_OSI
    t = small_iv_tree
    return iv_OSI(t["root"] + "/iv-osi", t["iv_models"], pre_model_dir=t["pre_model_dir"], threshold=0.0)


def test_posteriors_match_oracle(iv_osi, small_iv_tree):
    from fakebob_b200.engine import to_audio_list
    from oracle import kaldi_feats as kf
    system = small_iv_tree["system"]
    w = to_audio_list([make_audio(61, 0)])[0]
    iv_osi._engine.extract_ivectors([w])
    gsel, post = iv_osi._engine.posteriors()
    X = kf.voiced_features(w)
    og, op = system.posteriors(X)
    assert gsel.shape == og.shape
    # selection: identical except possibly around the 20th/21st boundary
    same_rows = sum(set(a) == set(b) for a, b in zip(gsel, og))
    assert same_rows >= 0.98 * len(og)
    bad = 0
    for t in range(len(og)):
        d_g = {int(c): float(p) for c, p in zip(gsel[t], post[t]) if p != 0}
        d_o = {int(c): float(p) for c, p in zip(og[t], op[t]) if p != 0}
        if set(d_g) != set(d_o):
            bad += 1                       # a posterior within rounding of min_post may be kept on one side only
            continue
        assert max(abs(d_g[c] - d_o[c]) for c in d_o) < TOL_POST
    assert bad <= 0.02 * len(og)


def test_ivectors_and_plda_scores_match_oracle(iv_osi, small_iv_tree):
    from fakebob_b200.engine import to_audio_list
    system = small_iv_tree["system"]
    lst = to_audio_list([make_audio(62, 0), make_audio(63, 1, n=24000), make_audio(64, 2, n=40000)])
    scores, ivs = iv_osi._engine.score_plda(lst, want_ivectors=True)
    ref_iv = np.stack([system.extract(w) for w in lst])
    assert np.abs(ivs - ref_iv).max() < TOL_IVEC
    order = np.argsort([m[0] for m in small_iv_tree["iv_models"]])
    enrolled = small_iv_tree["enrolled"][order]
    ref_scores = system.plda_scores(enrolled, ref_iv)
    assert scores.shape == ref_scores.shape == (3, 3)
    assert np.abs(scores - ref_scores).max() < TOL_PLDA
    # back-end alone (same i-vectors in): tight
    assert np.abs(system.plda_scores(enrolled, ivs) - scores).max() < 1e-6


def test_iv_wrappers_match_oracle_scorers(iv_osi, small_iv_tree):
    from fakebob_b200.ivector_PLDA_CSI import iv_CSI
    from fakebob_b200.ivector_PLDA_SV import iv_SV
    from oracle.scorers import OracleIvOSI, OracleIvSV
    t = small_iv_tree
    models = sorted(t["iv_models"], key=lambda m: m[0])
    enrolled = np.stack([np.asarray(__import__("fakebob_b200.kaldi_io", fromlist=["x"]).read_vector(m[2])) for m in models])
    ref = OracleIvOSI(t["system"], enrolled, [m[3] for m in models], [m[4] for m in models], threshold=0.0)
    batch = np.stack([make_audio(s, s % 3) for s in range(71, 74)], axis=1)
    got, want = iv_osi.score(batch), ref.score(batch)
    assert got.shape == want.shape == (3, 3)
    tol = TOL_PLDA / min(m[4] for m in models) + 1e-9
    assert np.abs(got - want).max() < tol
    assert iv_osi.spk_ids == sorted(iv_osi.spk_ids)
    one = iv_osi.score(batch[:, 0])
    assert one.shape == (3,)
    d, s = iv_osi.make_decisions(batch[:, 0])
    assert d in (-1, 0, 1, 2) and s.shape == (3,)
    csi = iv_CSI(t["root"] + "/iv-csi", t["iv_models"][::-1], pre_model_dir=t["pre_model_dir"])
    assert csi.spk_ids == iv_osi.spk_ids and np.allclose(csi.score(batch), got)
    sv = iv_SV(t["root"] + "/iv-sv", models[1], pre_model_dir=t["pre_model_dir"], threshold=0.0)
    rsv = OracleIvSV(t["system"], enrolled[1], models[1][3], models[1][4])
    g1, w1 = sv.score(batch), rsv.score(batch)
    assert g1.shape == (3,) and np.abs(g1 - w1).max() < TOL_PLDA / models[1][4] + 1e-9
    g2 = sv.score(batch[:, 0])
    assert np.isscalar(g2) or np.ndim(g2) == 0
    assert sv.make_decisions_value(1e9) == 1 and sv.make_decisions_value(-1e9) == -1


@pytest.mark.parametrize("task", ["SV", "OSI"])
def test_iv_attack_bit_exact_given_same_scores(small_iv_tree, iv_osi, task):
    from fakebob_b200.FAKEBOB import FakeBob
    from fakebob_b200.ivector_PLDA_SV import iv_SV
    from oracle.nes import OracleFakeBob
    t = small_iv_tree
    if task == "SV":
        model = iv_SV(t["root"] + "/iv-sv-nes", sorted(t["iv_models"], key=lambda m: m[0])[0], pre_model_dir=t["pre_model_dir"])
        kw = dict(threshold=float(model.score(make_audio(81, 0, n=16000))) + 0.5)
    else:
        model = iv_osi
        kw = dict(threshold=float(np.max(model.score(make_audio(81, 0, n=16000)))) + 0.5)
    audio = make_audio(81, 0, n=16000)
    hp = dict(max_iter=6, samples_per_draw=8)
    np.random.seed(3)
    fb = FakeBob(task, "untargeted", model, rng="numpy", verbose=False, **hp)
    adv_g, flag_g = fb.attack(audio.copy(), None, **kw)
    np.random.seed(3)
    ob = OracleFakeBob(task, "untargeted", model, **hp)
    adv_o, flag_o = ob.attack(audio.copy(), None, **kw)
    assert flag_g == flag_o and fb.iters_done == len(ob.log)
    assert np.array_equal(adv_g, adv_o)
    assert np.array_equal(fb.log[:, 1], np.array([float(np.asarray(r[1]).reshape(-1)[0]) for r in ob.log]))


def test_ivector_path_is_bitwise_reproducible(iv_osi):
    """The TMA-staged kernels (posteriors, quad) synchronise generic-proxy reads of bulk-copied shared memory and their
    partial-sum ring through mbarriers only; the slice partials are combined in a fixed order.  Same batch in -> the same
    bits out, every time (a timing-dependent hazard would show up as a flipped low bit in some repetition).  The batch
    mixes near-duplicate audios (shared Gaussian selections, like an NES batch) with unrelated ones."""
    from fakebob_b200.engine import to_audio_list
    base = to_audio_list([make_audio(71, 0)])[0]                 # int16
    rng = np.random.RandomState(5)
    near = [np.clip(base.astype(np.int32) + rng.randint(-3, 4, size=base.shape), -32768, 32767).astype(np.int16) for _ in range(9)]
    lst = to_audio_list(near + [make_audio(72, 1, n=24000), make_audio(73, 2, n=40000)])
    ref = None
    for rep in range(12):
        scores, ivs = iv_osi._engine.score_plda(lst, want_ivectors=True)
        gsel, post = iv_osi._engine.posteriors()
        cur = (scores.copy(), ivs.copy(), gsel.copy(), post.copy())
        if ref is None:
            ref = cur
            assert np.isfinite(scores).all() and np.isfinite(ivs).all()
        else:
            for a, b in zip(ref, cur):
                assert np.array_equal(a, b), "repetition %d differs" % rep


def test_register_staged_fallback_kernels_agree_with_the_tma_staged_ones(small_iv_tree):
    """The TMA-staged posterior / quad / lin kernels replaced register-staged ones that remain as the fallback for extractor
    sizes without 16-byte aligned rows (and as diagnostics).  Same batch through both sets, each in
    its own process (the switches are read once): same scores and i-vectors up to float rounding of the posteriors."""
    import json
    import os
    import subprocess
    import sys
    t = small_iv_tree
    code = r'''
import json, sys, os
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np
from conftest import test_audio as make_audio
from fakebob_b200.engine import to_audio_list, IvectorEngine
eng = IvectorEngine(%(pre)r)
eng.set_enrolled(np.load(%(enr)r))
lst = to_audio_list([make_audio(81, 0), make_audio(82, 1, n=24000), make_audio(83, 2, n=40000)])
scores, ivs = eng.score_plda(lst, want_ivectors=True)
print(json.dumps({"scores": np.asarray(scores).tolist(), "ivs": np.asarray(ivs).tolist()}))
'''
    enr = os.path.join(t["root"], "enrolled_for_fallback_test.npy")
    np.save(enr, np.asarray(t["enrolled"], dtype=np.float32))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = code % {"root": root, "pre": t["pre_model_dir"], "enr": enr}
    out = {}
    for name, extra in (("tma", {}), ("plain", {"FB_IV_PLAIN_POST": "1", "FB_IV_PLAIN_QUAD": "1", "FB_IV_PLAIN_LIN": "1"})):
        env = dict(os.environ)
        env.update(extra)
        r = subprocess.run([sys.executable, "-c", src], env=env, capture_output=True, text=True, timeout=240)
        assert r.returncode == 0, r.stderr[-2000:]
        out[name] = json.loads(r.stdout.strip().splitlines()[-1])
    a, b = np.array(out["tma"]["ivs"]), np.array(out["plain"]["ivs"])
    assert np.abs(a - b).max() < 2e-3 * max(1.0, np.abs(b).max())
    sa, sb = np.array(out["tma"]["scores"]), np.array(out["plain"]["scores"])
    assert np.abs(sa - sb).max() < 5e-3


def test_large_batch_rows_equal_single_utterance_results(iv_osi):
    """70 utterances in one call (three 32-utterance chunks in the quad / lin kernels, nine 8-row groups per frame index in
    the posterior kernel, active lists that are unions over a chunk) against the same utterances scored alone: an
    utterance's result must not depend on what else is in the batch."""
    from fakebob_b200.engine import to_audio_list
    lst = to_audio_list([make_audio(200 + i, i % 5, n=12000 + 800 * (i % 7)) for i in range(70)])
    scores, ivs = iv_osi._engine.score_plda(lst, want_ivectors=True)
    assert np.isfinite(scores).all() and np.isfinite(ivs).all()
    for i in (0, 7, 31, 32, 33, 63, 64, 69):
        s1, v1 = iv_osi._engine.score_plda([lst[i]], want_ivectors=True)
        assert np.abs(v1[0] - ivs[i]).max() < 1e-4 * max(1.0, np.abs(v1[0]).max()), i
        assert np.abs(s1[0] - scores[i]).max() < 1e-3, i
