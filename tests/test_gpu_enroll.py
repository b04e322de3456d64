"""GPU parity tests for enrolment: fb_map_adapt_host and fakebob_b200.build_spk_models against the
oracle restatement of gmm-global-acc-stats + gmm-global-est-map --update-flags=m (build_spk_models.py:202-219)."""
import os
import pickle

import numpy as np
import pytest

from conftest import load_oracle_gmm, test_audio as make_audio

pytestmark = pytest.mark.gpu

TOL_OCC = 2e-2         # absolute, component occupancies (sum to the number of voiced frames, ~190)
TOL_MIV = 2e-3         # absolute, means_invvars of magnitude up to ~30 (posterior differences of ~1e-4 times x / var)
TOL_GCONST = 2e-3      # absolute, gconsts of magnitude ~150


def test_map_adapt_matches_oracle(small_tree):
    from fakebob_b200 import kaldi_io
    from fakebob_b200.engine import GmmEngine, to_audio_list
    from oracle import kaldi_feats as kf
    ubm = kaldi_io.read_diag_gmm(small_tree["ubm"])
    eng = GmmEngine([ubm])
    w = to_audio_list([make_audio(41, 1)])
    got = eng.map_adapt(w, mean_tau=10.0)
    eng.close()
    ref = load_oracle_gmm(small_tree["ubm"]).map_adapt_means(kf.voiced_features(w[0]), tau=10.0)
    assert abs(got["occupancy"].sum() - ref.occupancy.sum()) < 1e-2            # both = number of voiced frames
    assert np.abs(got["occupancy"] - ref.occupancy).max() < TOL_OCC
    assert np.abs(got["means_invvars"] - ref.means_invvars).max() < TOL_MIV
    assert np.abs(got["gconsts"] - ref.gconsts).max() < TOL_GCONST
    assert np.array_equal(got["inv_vars"], ref.inv_vars) and np.array_equal(got["weights"], ref.weights)
    # the adapted means moved toward the data where there is occupancy, nowhere else
    moved = np.abs(got["means_invvars"] - ubm["means_invvars"]).max(axis=1)
    assert moved[got["occupancy"] < 1e-6].max(initial=0.0) < 1e-4
    assert moved[got["occupancy"] > 1.0].min() > 0


def test_pooled_utterances_and_tau(small_tree):
    """Several utterances are pooled like one feature archive; tau -> infinity keeps the UBM."""
    from fakebob_b200 import kaldi_io
    from fakebob_b200.engine import GmmEngine, to_audio_list
    ubm = kaldi_io.read_diag_gmm(small_tree["ubm"])
    eng = GmmEngine([ubm])
    a, b = to_audio_list([make_audio(42, 0), make_audio(43, 0, n=20000)])
    ga, gb, gab = eng.map_adapt([a]), eng.map_adapt([b]), eng.map_adapt([a, b])
    assert np.abs(gab["occupancy"] - (ga["occupancy"] + gb["occupancy"])).max() < 1e-9 * 1e3
    frozen = eng.map_adapt([a], mean_tau=1e12)
    assert np.abs(frozen["means_invvars"] - ubm["means_invvars"]).max() < 1e-4
    assert np.abs(frozen["gconsts"] - ubm["gconsts"]).max() < 1e-3
    eng.close()


def test_build_spk_models_round_trip(small_iv_tree, tmp_path):
    """enroll_gmm / enroll_ivector write the reference's file formats; the task wrappers load them and score like the oracle."""
    from fakebob_b200 import build_spk_models as bsm, kaldi_io
    from fakebob_b200.gmm_ubm_CSI import gmm_CSI
    from fakebob_b200.ivector_PLDA_CSI import iv_CSI
    from oracle import kaldi_feats as kf
    from oracle.scorers import OracleGmmCSI
    t = small_iv_tree
    spk = ["2001", "2002"]
    utt = ["2001-a-1", "2002-b-7"]
    enrol = [make_audio(51, 0), make_audio(52, 1)]
    cohort = [make_audio(53 + i, 2) for i in range(3)]
    model_dir = str(tmp_path / "model")
    gm = bsm.enroll_gmm(enrol, spk, utt, cohort, t["pre_model_dir"], model_dir)
    assert [m[0] for m in gm] == spk and [m[1] for m in gm] == utt
    for m in gm:
        assert os.path.isabs(m[2]) and m[2].endswith(m[0] + "-identity.gmm") and m[4] > 0
        with open(os.path.join(model_dir, m[0] + ".gmm"), "rb") as f:
            assert pickle.load(f) == m
    # oracle enrolment + oracle z-norm on the same audio
    from fakebob_b200.engine import to_audio_list
    ubm = load_oracle_gmm(t["ubm"])
    ident = [ubm.map_adapt_means(kf.voiced_features(w), tau=10.0) for w in to_audio_list(enrol)]
    zs = np.array([[float(g.avg_loglike(kf.voiced_features(w))) for g in ident] for w in to_audio_list(cohort)])
    assert np.abs(np.array([m[3] for m in gm]) - zs.mean(axis=0)).max() < 1e-3
    assert np.abs(np.array([m[4] for m in gm]) - zs.std(axis=0)).max() < 1e-3
    csi = gmm_CSI(str(tmp_path / "grp"), gm, pre_model_dir=t["pre_model_dir"])
    ref = OracleGmmCSI(ident, [m[3] for m in gm], [m[4] for m in gm])
    probe = make_audio(57, 0)
    assert np.abs(csi.score(probe) - ref.score(probe)).max() < 5e-3        # z-normalised: divided by a std of ~0.3
    # i-vector side
    im = bsm.enroll_ivector(enrol, spk, utt, cohort, t["pre_model_dir"], model_dir)
    system = t["system"]
    ref_iv = np.stack([system.extract(w) for w in to_audio_list(enrol)])
    got_iv = np.stack([kaldi_io.read_vector(m[2]) for m in im])
    assert np.abs(got_iv - ref_iv).max() < 5e-3
    ref_sc = system.plda_scores(got_iv, np.stack([system.extract(w) for w in to_audio_list(cohort)]))
    assert np.abs(np.array([m[3] for m in im]) - ref_sc.mean(axis=0)).max() < 2e-2
    assert np.abs(np.array([m[4] for m in im]) - ref_sc.std(axis=0)).max() < 2e-2
    iv = iv_CSI(str(tmp_path / "grp-iv"), im, pre_model_dir=t["pre_model_dir"])
    assert iv.spk_ids == spk and iv.score(probe).shape == (2,)
