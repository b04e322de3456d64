"""GPU edge cases of the scoring and attack paths: long utterances (global-scratch front-end path), very short ones, odd and
large samples_per_draw, empty batches, int16 pass-through, argument errors."""
import numpy as np
import pytest

from conftest import test_audio as make_audio

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def osi(small_tree):
    from fakebob_b200.gmm_ubm_OSI import gmm_OSI
    return gmm_OSI(small_tree["root"] + "/grp-edges", small_tree["models"], small_tree["ubm"],
                   pre_model_dir=small_tree["pre_model_dir"], threshold=0.1)


def test_long_utterance_uses_global_scratch_and_matches_oracle(osi, small_oracle_models):
    """40 s = 4000 frames exceeds the shared-memory budget of feats_kernel (prefix sums + raw features), so the kernel takes
    its global-scratch path; scores must still match the oracle and the shared-memory path on a shorter copy."""
    from oracle.scorers import OracleGmmOSI
    ubm, spk = small_oracle_models
    long_audio = np.concatenate([make_audio(200 + i, i % 3, n=160000) for i in range(4)])
    assert long_audio.shape[0] == 640000
    got = osi.score(long_audio)
    want = OracleGmmOSI(ubm, spk).score(long_audio)
    assert got.shape == (3,) and np.abs(got - want).max() < 5e-4
    # mixed batch: the long utterance next to a short one
    both = osi.score([long_audio, make_audio(5, 1)])
    assert np.abs(both[0] - got).max() < 2e-4


def test_short_utterance_and_int16_passthrough(osi, small_oracle_models):
    from fakebob_b200.engine import to_audio_list
    from oracle.scorers import OracleGmmOSI
    ubm, spk = small_oracle_models
    a = make_audio(77, 2, n=4000)                       # 0.25 s: 25 frames, reflected edges dominate
    got = osi.score(a)
    want = OracleGmmOSI(ubm, spk).score(a)
    assert np.abs(got - want).max() < 5e-4
    as_int = to_audio_list([a])[0]
    assert as_int.dtype == np.int16
    assert np.array_equal(osi.score(as_int), got)       # int16 input is not rescaled (gmm_ubm_OSI.py:83-85)
    assert np.array_equal(osi.score(a[:, None]), got) and np.array_equal(osi.score(a[None, :]), got)


def test_empty_batch_and_bad_arguments(osi):
    from fakebob_b200._lib import FakebobLibraryError
    assert osi._engine.score_avg_ll([]).shape == (0, 4)
    with pytest.raises((FakebobLibraryError, ValueError)):
        osi._engine.score_avg_ll([np.zeros(10, dtype=np.int16)])          # shorter than half a frame shift
    with pytest.raises(ValueError):
        osi.score(np.zeros((2, 3, 4)))


@pytest.mark.parametrize("S", [2, 7, 256])
def test_samples_per_draw_extremes(osi, S):
    """One antithetic pair; odd S uses S-1 samples (FAKEBOB.py:234-235); S = 256 is configs[3]'s draw.  (S = 1 has no pair:
    the reference takes the mean of an empty slice -> NaN; here fb_nes_init rejects it.)"""
    from fakebob_b200.FAKEBOB import FakeBob
    from oracle.nes import OracleFakeBob, PhiloxNoise
    audio = make_audio(31, 0, n=16000)
    hp = dict(max_iter=3, samples_per_draw=S)
    fb = FakeBob("OSI", "untargeted", osi, seed=11, verbose=False, **hp)
    adv, flag = fb.attack(audio, None, threshold=1e3)
    ob = OracleFakeBob("OSI", "untargeted", osi, noise_fn=PhiloxNoise(11), **hp)
    adv_o, flag_o = ob.attack(audio, None, threshold=1e3)
    assert flag == flag_o == -1 and fb.iters_done == 3
    assert np.mean(adv == adv_o) > 0.999
    if S == 2:
        from fakebob_b200._lib import FakebobLibraryError
        with pytest.raises(FakebobLibraryError):
            FakeBob("OSI", "untargeted", osi, max_iter=1, samples_per_draw=1, verbose=False).attack(audio, None, threshold=1e3)


def test_attack_many_shards_utterances(osi):
    from fakebob_b200.FAKEBOB import FakeBob
    from fakebob_b200.sharding import attack_many, utterance_range
    audios = [make_audio(40 + i, i % 3, n=16000) for i in range(5)]
    assert [utterance_range(5, r, 2) for r in range(2)] == [(0, 2), (2, 5)]
    mk = lambda: FakeBob("OSI", "untargeted", osi, max_iter=2, samples_per_draw=4, seed=5, verbose=False)
    parts = [attack_many(mk, audios, rank=r, world=2, threshold=1e3) for r in range(2)]
    assert sorted(list(parts[0]) + list(parts[1])) == [0, 1, 2, 3, 4]
    whole = attack_many(mk, audios, threshold=1e3)
    for p in parts:
        for u, (adv, flag) in p.items():
            assert flag == whole[u][1] and np.array_equal(adv, whole[u][0])
