"""Parity at the sizes bench.py measures (BASELINE.json configs[1] "C2" and configs[2] "C3"), through the C-ABI, against
oracle outputs committed in tests/golden/fullsize_c2.npz / fullsize_c3.npz (tests/golden/make_golden_fullsize.py; the
oracle needs minutes and 6 GB at these sizes, so it ran once in the CPU container).  The seeded synthetic model trees are
rebuilt here (tests/fullsize_util.py) and their parameter checksums are compared with the generator's.

Tolerances (absolute): MFCC / features 2e-3, per-frame log-likelihood 2e-3 (three-term contraction), average
log-likelihood 5e-4, scores 5e-4; with the default one-term difference contraction per-frame 1e-2 and scores 5e-4 (measured
deviation is reported by the test); posteriors 2e-3, gamma 5e-3, X / lin / quad relative 1e-4 of their scale, raw i-vector
5e-3, PLDA log-likelihood ratio 5e-2 (|LLR| ~ 200), z-normed score 5e-2 / z_std.
Also: size-independent properties (batch independence, permutation invariance, ragged batches, determinism, the
L-infinity box of the update)."""
import os

import numpy as np
import pytest

import fullsize_util as fu

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_MFCC = 2e-3
TOL_FEAT = 2e-3
TOL_FRAME_LL = 2e-3
TOL_FRAME_LL_DELTA1 = 1e-2
TOL_AVG_LL = 5e-4
TOL_SCORE = 5e-4
TOL_POST = 2e-3
TOL_IVEC = 5e-3
TOL_PLDA = 5e-2


@pytest.fixture(scope="module")
def c2(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("c2_tree"))
    tree = fu.build_c2_tree(root)
    gold = np.load(os.path.join(GOLD, "fullsize_c2.npz"))
    assert fu.checksums_close(fu.gmm_checksums(tree), gold["checksums"], rtol=1e-5), \
        "the synthetic C2 tree built here differs from the one the fixture was generated with"
    tree["gold"] = gold
    tree["paths"] = [tree["ubm"]] + [m[2] for m in tree["models"]]
    tree["waves"] = [np.ascontiguousarray(gold["wave%d" % u]) for u in range(fu.N_UTTS)]
    return tree


def test_c2_stages_match_fixture(c2):
    """Front-end and 2048-mixture GMM kernel, three-term contraction, 3 utterances x 6 models, stage by stage."""
    from fakebob_b200.engine import GmmEngine
    g = c2["gold"]
    eng = GmmEngine.from_files(c2["paths"], delta_terms=3)
    eng.set_debug(True)
    avg = eng.score_avg_ll(c2["waves"])
    st = eng.last_stages()
    f0 = r0 = 0
    worst = {"mfcc": 0.0, "feats": 0.0, "frame_ll": 0.0, "avg_ll": 0.0}
    for u in range(fu.N_UTTS):
        mf, vad, fl = g["mfcc%d" % u], g["vad%d" % u] != 0, g["frame_ll%d" % u]
        T, Tv = mf.shape[0], int(vad.sum())
        assert st["frames"][u] == T
        worst["mfcc"] = max(worst["mfcc"], float(np.abs(st["mfcc"][f0:f0 + T] - mf).max()))
        got_vad = st["vad"][f0:f0 + T] >= 0
        # a VAD decision may only differ where C0 is within float rounding of the threshold; then the voiced rows do not
        # line up frame by frame and only the utterance averages are compared (looser: one frame of ~450 differs)
        assert (got_vad != vad).sum() <= 1
        if np.array_equal(got_vad, vad):
            assert st["voiced"][u] == Tv
            if u == 0:
                worst["feats"] = float(np.abs(st["feats"][r0:r0 + Tv] - g["feats0"]).max())
            worst["frame_ll"] = max(worst["frame_ll"], float(np.abs(st["frame_ll"][:, r0:r0 + Tv] - fl).max()))
            worst["avg_ll"] = max(worst["avg_ll"], float(np.abs(avg[u] - g["avg_ll%d" % u]).max()))
        else:
            assert np.abs(avg[u] - g["avg_ll%d" % u]).max() < 0.1
        f0 += T
        r0 += int(st["voiced"][u])
    print("C2 stage deviations vs oracle fixture:", worst)
    assert worst["mfcc"] < TOL_MFCC and worst["feats"] < TOL_FEAT
    assert worst["frame_ll"] < TOL_FRAME_LL and worst["avg_ll"] < TOL_AVG_LL
    eng.close()


def test_c2_scores_default_contraction(c2):
    """The shipped configuration (automatic difference terms) through the reference-named wrappers: OSI, CSI, SV."""
    from fakebob_b200.gmm_ubm_CSI import gmm_CSI
    from fakebob_b200.gmm_ubm_OSI import gmm_OSI
    from fakebob_b200.gmm_ubm_SV import gmm_SV
    g = c2["gold"]
    avg = np.stack([g["avg_ll%d" % u] for u in range(fu.N_UTTS)])
    osi = gmm_OSI(c2["root"] + "/grp-osi", c2["models"], c2["ubm"], pre_model_dir=c2["pre_model_dir"])
    info = osi._engine.gmm_info()
    got = osi.score(c2["waves"])
    want = avg[:, 1:] - avg[:, :1]
    dev = float(np.abs(got - want).max())
    fl = osi._engine.last_stages()["frame_ll"]
    fdev = max(float(np.abs(fl[:, a:b] - g["frame_ll%d" % u]).max())
               for u, (a, b) in enumerate(zip(np.cumsum([0] + [int((g["vad%d" % v] != 0).sum()) for v in range(fu.N_UTTS)])[:-1],
                                              np.cumsum([int((g["vad%d" % v] != 0).sum()) for v in range(fu.N_UTTS)]))))
    print("C2 default contraction: %s, OSI score deviation vs oracle %.2e, frame LL deviation %.2e" % (info, dev, fdev))
    assert info["shared_variances"]
    assert dev < TOL_SCORE and fdev < TOL_FRAME_LL_DELTA1
    dec, sc = osi.make_decisions(c2["waves"])
    assert list(dec) == list(np.argmax(want, axis=1)) or np.max(want) < osi.threshold
    csi = gmm_CSI(c2["root"] + "/grp-csi", c2["models"], pre_model_dir=c2["pre_model_dir"])
    want_csi = (avg[:, 1:] - g["z_means"]) / g["z_stds"]
    assert np.abs(csi.score(c2["waves"]) - want_csi).max() < TOL_SCORE / float(np.min(g["z_stds"])) + 1e-9
    sv = gmm_SV(c2["root"] + "/spk-sv", c2["models"][2], c2["ubm"], pre_model_dir=c2["pre_model_dir"])
    assert np.abs(sv.score(c2["waves"]) - want[:, 2]).max() < TOL_SCORE


def test_c2_nes_trajectory_matches_fixture(c2):
    """3 iterations of the S = 50 attack on the device vs the oracle loop + oracle scorer, same Philox stream."""
    from fakebob_b200.FAKEBOB import FakeBob
    from fakebob_b200.gmm_ubm_OSI import gmm_OSI
    g = c2["gold"]
    osi = gmm_OSI(c2["root"] + "/grp-nes", c2["models"], c2["ubm"], pre_model_dir=c2["pre_model_dir"])
    audio = c2["waves"][0].astype(np.float64) / 32768.0
    fb = FakeBob("OSI", "untargeted", osi, epsilon=0.002, max_iter=fu.NES_ITERS, samples_per_draw=fu.NES_S,
                 seed=fu.NES_SEED, verbose=False)
    adv, flag = fb.attack(audio, None, threshold=fu.NES_THRESHOLD)
    assert flag == int(g["nes_flag"][0]) and fb.iters_done == fu.NES_ITERS
    want_adv = c2["waves"][0].astype(np.int32) + g["nes_adver_delta"].astype(np.int32)
    agree = float(np.mean(adv[:, 0].astype(np.int32) == want_adv))
    loss_dev = float(np.abs(fb.log[:, 1] - g["nes_adver_loss"]).max())
    print("C2 NES trajectory: adversarial samples identical %.5f, adver_loss deviation per iteration %s"
          % (agree, np.abs(fb.log[:, 1] - g["nes_adver_loss"])))
    assert np.abs(fb.log[:, 0] - g["nes_distance"]).max() < 1e-12
    # iteration 0 scores the clean audio: score tolerance.  Later iterations follow the sign of a gradient estimated from
    # loss differences of ~1e-3 between the 50 samples, so a 1e-4 score deviation flips the step of a few per cent of the
    # samples and the trajectories drift apart (the oracle itself moves as much when its BLAS changes summation order);
    # the loss must stay close and the adversarial audios overwhelmingly identical.
    assert abs(fb.log[0, 1] - g["nes_adver_loss"][0]) < TOL_SCORE and np.abs(fb.log[0, 4:] - g["nes_scores"][0]).max() < TOL_SCORE
    assert loss_dev < 5e-2
    assert agree > 0.95


def test_c2_batch_independence_permutation_ragged(c2):
    from fakebob_b200 import synth
    from fakebob_b200.engine import GmmEngine
    eng = GmmEngine.from_files(c2["paths"])
    batch = [synth.to_int16(synth.synth_utterance(300 + i, i % 7, fu.N_SAMPLES)) for i in range(fu.NES_S + 1)]
    full = eng.score_avg_ll(batch)
    assert full.shape == (fu.NES_S + 1, 6) and np.isfinite(full).all()
    rows = eng.voiced_rows()
    assert 0.5 * (fu.NES_S + 1) * 500 < rows <= (fu.NES_S + 1) * 500
    for i in (0, 17, fu.NES_S):
        alone = eng.score_avg_ll([batch[i]])
        assert np.abs(alone[0] - full[i]).max() < 2e-4          # same frames, different tiles / segment cuts
    rev = eng.score_avg_ll(batch[::-1])
    assert np.abs(rev[::-1] - full).max() < 2e-4
    # ragged: different lengths in one batch (loadData passes such lists, attackMain.py:128)
    ragged = [batch[0][:31234], batch[1], batch[2][:16000]]
    r = eng.score_avg_ll(ragged)
    assert np.abs(r[1] - full[1]).max() < 2e-4
    assert np.abs(eng.score_avg_ll([ragged[0]])[0] - r[0]).max() < 2e-4
    eng.close()


def test_c2_nes_deterministic_and_boxed(c2):
    from fakebob_b200 import synth
    from fakebob_b200.FAKEBOB import FakeBob
    from fakebob_b200.gmm_ubm_OSI import gmm_OSI
    osi = gmm_OSI(c2["root"] + "/grp-det", c2["models"], c2["ubm"], pre_model_dir=c2["pre_model_dir"])
    audio = synth.synth_utterance(400, 3, fu.N_SAMPLES)
    eps = 0.002
    outs = []
    for _ in range(2):
        fb = FakeBob("OSI", "untargeted", osi, epsilon=eps, max_iter=4, samples_per_draw=fu.NES_S, seed=99, verbose=False)
        adv, flag = fb.attack(audio.copy(), None, threshold=1e3)
        outs.append((adv.copy(), flag, fb.log.copy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1] == -1
    assert np.array_equal(outs[0][2], outs[1][2])                # distance / loss / lr / scores per iteration
    adv = outs[0][0]
    assert adv.dtype == np.int16 and adv.shape == (fu.N_SAMPLES, 1)
    # L-infinity box of FAKEBOB.py:202-203 in int16 units (truncation adds < 1)
    assert np.abs(adv[:, 0].astype(np.int64) - (audio * 32768).astype(np.int64)).max() <= int(eps * 32768) + 1
    log = outs[0][2]
    assert log.shape[0] == 4 and np.all(log[:, 0] <= eps + 1e-12)


# ------------------------------------------------------------------------------------------------ C3
@pytest.fixture(scope="module")
def c3(c2):
    from fakebob_b200.engine import IvectorEngine
    gold = np.load(os.path.join(GOLD, "fullsize_c3.npz"))
    fu.build_c3_params(c2["root"], c2)
    assert fu.checksums_close(fu.iv_checksums(c2["pre_model_dir"]), gold["checksums"], rtol=1e-5), \
        "the synthetic C3 parameters built here differ from the ones the fixture was generated with"
    eng = IvectorEngine(c2["pre_model_dir"])
    eng.set_enrolled(gold["enrolled"])
    return {"gold": gold, "eng": eng, "waves": c2["waves"], "pre_model_dir": c2["pre_model_dir"], "root": c2["root"]}


def test_c3_every_stage_matches_fixture(c3):
    """R = 400, L = 200, C = 2048: Gaussian selection, pruned posteriors, gamma / X, lin, quad, i-vector, PLDA LLR."""
    g, eng = c3["gold"], c3["eng"]
    scores, ivs = eng.score_plda(c3["waves"], want_ivectors=True)
    gsel, post = eng.posteriors()
    R = eng.R
    r0 = 0
    worst = {"post": 0.0, "gamma": 0.0, "X_rel": 0.0, "lin_rel": 0.0, "quad_rel": 0.0, "ivector": 0.0, "llr": 0.0}
    bad_total = 0
    for u in range(fu.N_UTTS):
        og, op = g["gsel%d" % u].astype(np.int64), g["post%d" % u]
        Tv = og.shape[0]
        gg, gp = gsel[r0:r0 + Tv], post[r0:r0 + Tv]
        same_sel = sum(set(a) == set(b) for a, b in zip(gg, og))
        assert same_sel >= 0.98 * Tv                               # ties around the 20th / 21st component may differ
        bad = 0
        for t in range(Tv):
            d_g = {int(c): float(p) for c, p in zip(gg[t], gp[t]) if p != 0}
            d_o = {int(c): float(p) for c, p in zip(og[t], op[t]) if p != 0}
            if set(d_g) != set(d_o):
                bad += 1                                           # a posterior within rounding of min_post kept on one side only
                continue
            worst["post"] = max(worst["post"], max(abs(d_g[c] - d_o[c]) for c in d_o))
        assert bad <= 0.02 * Tv
        bad_total += bad
        st = eng.stats(u)
        worst["gamma"] = max(worst["gamma"], float(np.abs(st["gamma"] - g["gamma%d" % u]).max()))
        xs = np.abs(g["xsum%d" % u]).max()
        worst["X_rel"] = max(worst["X_rel"], float(np.abs(st["X"].sum(axis=1) - g["xsum%d" % u]).max() / xs))
        if u == 0:
            worst["X_rel"] = max(worst["X_rel"], float(np.abs(st["X"] - g["X0"]).max() / np.abs(g["X0"]).max()))
            worst["quad_rel"] = max(worst["quad_rel"], float(np.abs(st["quad"] - g["quad0_tril"]).max() / np.abs(g["quad0_tril"]).max()))
        worst["lin_rel"] = max(worst["lin_rel"], float(np.abs(st["lin"] - g["lin%d" % u]).max() / np.abs(g["lin%d" % u]).max()))
        diag = st["quad"][np.cumsum(np.arange(1, R + 1)) - 1]
        worst["quad_rel"] = max(worst["quad_rel"], float(np.abs(diag - g["quad_diag%d" % u]).max() / np.abs(g["quad_diag%d" % u]).max()))
        row0 = st["quad"][np.arange(R) * (np.arange(R) + 1) // 2]          # column 0 of the lower triangle = row 0 of the matrix
        worst["quad_rel"] = max(worst["quad_rel"], float(np.abs(row0 - g["quad_row0_%d" % u]).max() / np.abs(g["quad_row0_%d" % u]).max()))
        worst["ivector"] = max(worst["ivector"], float(np.abs(ivs[u] - g["ivector%d" % u]).max()))
        worst["llr"] = max(worst["llr"], float(np.abs(scores[u] - g["llr%d" % u]).max()))
        r0 += Tv
    print("C3 stage deviations vs oracle fixture:", worst, "frames with a posterior pruned on one side only:", bad_total)
    # every frame whose pruning decision differs (posterior within rounding of min_post = 0.025) moves <= 0.05 of occupancy
    slack = float(bad_total)
    assert worst["post"] < TOL_POST and worst["gamma"] < 5e-3 + 0.08 * slack
    assert worst["X_rel"] < 1e-4 + 2e-3 * slack and worst["lin_rel"] < 1e-4 + 2e-3 * slack and worst["quad_rel"] < 1e-4 + 2e-3 * slack
    assert worst["ivector"] < TOL_IVEC * (1 + slack) and worst["llr"] < TOL_PLDA * (1 + slack)


def test_c3_wrappers_match_fixture(c3):
    """iv_SV / iv_OSI / iv_CSI with the enrolled identities written in the reference's file formats."""
    import pickle
    from fakebob_b200 import kaldi_io
    from fakebob_b200.ivector_PLDA_CSI import iv_CSI
    from fakebob_b200.ivector_PLDA_OSI import iv_OSI
    from fakebob_b200.ivector_PLDA_SV import iv_SV
    g = c3["gold"]
    spk_ids = [str(s) for s in g["spk_ids"]]
    ark = os.path.join(c3["root"], "enrolled-fixture.ark")
    targets = kaldi_io.write_text_vector_ark(ark, [(s + "-enroll", g["enrolled"][k]) for k, s in enumerate(spk_ids)])
    models = [[s, s + "-enroll", targets[s + "-enroll"], float(g["z_means"][k]), float(g["z_stds"][k])] for k, s in enumerate(spk_ids)]
    want = np.stack([g["score%d" % u] for u in range(fu.N_UTTS)])
    tol = TOL_PLDA / float(np.min(g["z_stds"])) + 1e-6
    osi = iv_OSI(c3["root"] + "/iv-osi", models, pre_model_dir=c3["pre_model_dir"], threshold=0.0)
    got = osi.score(c3["waves"])
    print("C3 z-normed score deviation vs oracle fixture: %.3e (tolerance %.3e)" % (np.abs(got - want).max(), tol))
    assert got.shape == want.shape and np.abs(got - want).max() < tol
    csi = iv_CSI(c3["root"] + "/iv-csi", models[::-1], pre_model_dir=c3["pre_model_dir"])
    assert np.allclose(csi.score(c3["waves"]), got)
    sv = iv_SV(c3["root"] + "/iv-sv", models[1], pre_model_dir=c3["pre_model_dir"], threshold=0.0)
    s1 = sv.score(c3["waves"])
    assert s1.shape == (fu.N_UTTS,) and np.abs(s1 - want[:, 1]).max() < tol
    assert pickle.dumps(models)                                    # the 5-list layout of build_spk_models.py:152


def test_c3_long_utterance_statistics_in_chunks(c3):
    """An utterance longer than the shared-memory bucket capacity of the statistics kernel (~830 frames) is accumulated in
    frame chunks; its statistics must equal the sum over the same frames scored as separate utterances only in gamma
    (X differs because CMN windows differ), so compare against the device's own single-pass result on a length that fits."""
    from fakebob_b200 import synth
    eng = c3["eng"]
    long_w = synth.to_int16(np.concatenate([synth.synth_utterance(900 + i, i, 80000) for i in range(3)]))   # 15 s = 1500 frames
    iv_long = eng.extract_ivectors([long_w])
    st = eng.stats(0)
    assert np.isfinite(iv_long).all()
    gsel, post = eng.posteriors()
    # gamma from the posteriors directly (host) vs the chunked kernel
    gam = np.zeros(st["gamma"].shape[0])
    np.add.at(gam, gsel.reshape(-1), post.reshape(-1).astype(np.float64))
    assert np.abs(gam - st["gamma"]).max() < 1e-6 * max(1.0, gam.max())
    # batching a long and a short utterance leaves the short one's i-vector unchanged
    both = eng.extract_ivectors([long_w, c3["waves"][0]])
    alone = eng.extract_ivectors([c3["waves"][0]])
    assert np.abs(both[1] - alone[0]).max() < 1e-4 and np.abs(both[0] - iv_long[0]).max() < 1e-4
