"""CPU tests of oracle/kaldi_nonideal.py: the switches for the non-ideal effects of the reference's real Kaldi path
(SURVEY.md A.9): libc generators behind Kaldi's dither (checked against this machine's C library), the CompressedMatrix
speech-feature codec (structural properties + hand-computed values), 7-significant-digit text round trips."""
import ctypes
import ctypes.util

import numpy as np
import pytest

from oracle import kaldi_feats as kf
from oracle import kaldi_nonideal as kn


def _libc():
    name = ctypes.util.find_library("c")
    if not name:
        pytest.skip("no C library to compare with")
    return ctypes.CDLL(name)


def test_glibc_rand_matches_the_c_library():
    libc = _libc()
    libc.srand(1)
    want = [libc.rand() for _ in range(2000)]
    g = kn.GlibcRand(1)
    assert [g.rand() for _ in range(2000)] == want
    libc.srand(12345)
    g = kn.GlibcRand(12345)
    assert [g.rand() for _ in range(100)] == [libc.rand() for _ in range(100)]


def test_rand_r_matches_the_c_library():
    libc = _libc()
    libc.rand_r.argtypes = [ctypes.POINTER(ctypes.c_uint)]
    for seed in (1, 27437 + 1804289383, 0xFFFFFFFF):
        s = ctypes.c_uint(seed & 0xFFFFFFFF)
        want = [libc.rand_r(ctypes.byref(s)) for _ in range(500)]
        got, cur = [], seed & 0xFFFFFFFF
        for _ in range(500):
            v, cur = kn.rand_r(cur)
            got.append(v)
        assert got == want and cur == s.value
        assert kn.rand_r_stream(seed, 500).tolist() == want


def test_dither_is_unit_variance_noise_and_deterministic():
    fr = np.zeros((40, 400), dtype=np.float32)
    a = kn.dither_frames(fr, 1.0)
    b = kn.dither_frames(fr, 1.0)
    assert np.array_equal(a, b)                                   # a fresh process always sees the same stream (seed 1)
    assert abs(float(a.mean())) < 0.05 and abs(float(a.std()) - 1.0) < 0.05
    assert not np.array_equal(a[0], a[1])                         # a new RandomState per frame
    assert np.array_equal(kn.dither_frames(fr, 0.5), (0.5 * a).astype(np.float32))


def test_compressed_matrix_round_trip_properties():
    r = np.random.default_rng(3)
    M = (r.standard_normal((500, 24)) * np.linspace(1, 20, 24)[None, :] + np.linspace(-30, 30, 24)[None, :]).astype(np.float32)
    D = kn.compress_decompress(M)
    assert D.shape == M.shape and D.dtype == np.float32
    rng = float(M.max() - M.min())
    col_range = M.max(0) - M.min(0)
    # one byte per value, piecewise linear between the column's 0 / 25 / 75 / 100 % points: the error is a fraction of the
    # column's own spread (plus the 16-bit quantisation of the four points against the GLOBAL range)
    assert np.all(np.abs(D - M).max(0) <= col_range / 63.0 / 2 * 1.01 + rng * 2e-5)
    # order preserving within a column (the codec is monotone)
    for d in (0, 7, 23):
        o = np.argsort(M[:, d], kind="stable")
        assert np.all(np.diff(D[o, d]) >= -1e-6)
    # at most 256 distinct values per column; compressing the decoded matrix again changes (almost) nothing
    assert all(len(np.unique(D[:, d])) <= 256 for d in range(24))
    D2 = kn.compress_decompress(D)
    assert np.abs(D2 - D).max() <= col_range.max() / 63.0 / 2 * 1.01 + rng * 2e-5
    with pytest.raises(ValueError):
        kn.compress_decompress(M[:8])


def test_compressed_matrix_hand_computed_column():
    # a single column 0..15 (T = 16): min 0, range 15, p0 = 0, p25 = value at sorted[4] = 4, p75 = sorted[12] = 12, p100 = 15
    M = np.arange(16, dtype=np.float32)[:, None]
    D = kn.compress_decompress(M)[:, 0]
    u = lambda v: int(np.float32(v / 15.0) * np.float32(65535.0) + np.float32(0.499))
    f = lambda p: np.float32(0.0) + np.float32(15.0) * np.float32(1.52590218966964e-05) * np.float32(p)
    p0, p25, p75, p100 = f(u(0)), f(u(4)), f(u(12)), f(u(15))
    # value 2 sits halfway in the first segment: byte 32 -> p0 + (p25 - p0) * 32 / 64
    assert abs(D[2] - (p0 + (p25 - p0) * 0.5)) < 1e-5
    # value 8 sits halfway in the middle segment: byte 64 + 64
    assert abs(D[8] - (p25 + (p75 - p25) * 0.5)) < 1e-5
    assert abs(D[15] - p100) < 1e-5 and abs(D[0] - p0) < 1e-5


def test_seven_digit_text_round_trip():
    assert kn.round_sig7(-119.87654321) == -119.8765
    assert kn.round_sig7(0.000123456789) == 0.0001234568
    assert kn.round_sig7(np.array([1.0, 2.5e10, -3.14159265358979])).tolist() == [1.0, 2.5e10, -3.141593]
    v = np.array([0.123456789, -12.3456789], dtype=np.float32)
    assert kn.round_sig7_f32(v).dtype == np.float32 and np.allclose(kn.round_sig7_f32(v), [0.1234568, -12.34568], rtol=1e-7)


def test_nonideal_effects_move_scores_by_what_the_design_note_says():
    """Size of the three effects on one utterance's features / average log-likelihood (documents DESIGN.md section 3)."""
    from fakebob_b200 import synth
    from oracle.diag_gmm import DiagGmm
    wave = synth.to_int16(synth.synth_utterance(3, 1, 32000))
    cfg = kf.FeatConfig()
    m = kf.mfcc(wave, cfg)
    fr = kf.extract_frames(wave, cfg)
    assert fr.shape == (200, 400)
    md = kf.mfcc(wave, cfg, frames=kn.dither_frames(fr, 1.0))
    mc = kn.compress_decompress(m)
    assert 0 < np.abs(md - m).max() < 5.0 and np.median(np.abs(md - m)) < 0.05   # +-1 LSB of noise: only near-silent frames move
    assert 0 < np.abs(mc - m).max() < 0.5 * float(m.max() - m.min()) / 63.0 + 1e-3
    X = kf.select_voiced(kf.sliding_cmn(kf.add_deltas(m)), kf.compute_vad(m))
    r = np.random.default_rng(1)
    mu = X[r.integers(0, X.shape[0], 64)].astype(np.float64)
    g = DiagGmm.from_moments(np.full(64, 1 / 64), mu, np.tile(X.var(0), (64, 1)))
    ll = float(g.avg_loglike(X))
    assert abs(kn.round_sig7(ll) - ll) < 5e-5 * max(1.0, abs(ll) / 100.0) * 2
