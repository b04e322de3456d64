"""Oracle (TEST INFRASTRUCTURE): the three non-ideal effects of the reference's real Kaldi path, as switches.
PARITY UNPINNED against real Kaldi (see oracle/__init__.py): restated from the published upstream sources.

The reference runs Kaldi through shell scripts, so every score() call also goes through
  * dither ................ compute-mfcc-feats with Kaldi's default --dither=1.0 (conf/mfcc.conf does not override it,
                            gmm_ubm_kaldiHelper.py:138): feat/feature-window.cc Dither() adds RandGauss() * dither to every sample of
                            every frame, from a per-frame RandomState seeded by the process-wide rand() (base/kaldi-math.cc);
  * feature compression ... steps/make_mfcc.sh pipes through copy-feats --compress=true (gmm_ubm_kaldiHelper.py:138-147):
                            matrix/compressed-matrix.cc, speech-feature format (one byte per value, four percentiles per column);
  * 7-digit text .......... scores are written with `ark,t:` (gmm_ubm_kaldiHelper.py:204-208) and i-vectors with `ark,scp,t:`
                            (ivector_PLDA_kaldiHelper.py:202-211); Kaldi's text output stream has precision(7).
The device path runs with all three OFF by default; `oracle.kaldi_feats.FeatConfig.dither` and the
functions below let the oracle reproduce them, and fakebob_b200's `kaldi_exact` options enable compression and text rounding
on the device.  The dither stream depends on libc's rand() / rand_r() (glibc algorithms restated here and checked against
the C library in tests/test_oracle_nonideal.py) and on the order in which one Kaldi process meets the frames, so it is
exact only for a single-job (`nj=1`) run.
"""
import numpy as np

F32 = np.float32
RAND_MAX = 2147483647


# ------------------------------------------------------------------------------------------------ text precision
def round_sig7(x):
    """What survives `os << x` with precision(7) followed by reading the text back (float64 in, float64 out)."""
    a = np.asarray(x, dtype=np.float64)
    out = np.array([float("%.7g" % v) for v in a.reshape(-1)], dtype=np.float64).reshape(a.shape)
    return out if out.shape else float(out)


def round_sig7_f32(x):
    """The same for a Vector<BaseFloat> (enrolled / test i-vectors): float32 written as text, read back as float32."""
    a = np.asarray(x, dtype=F32)
    return np.array([F32(float("%.7g" % float(v))) for v in a.reshape(-1)], dtype=F32).reshape(a.shape)


# ------------------------------------------------------------------------------------------------ CompressedMatrix
def _float_to_u16(vmin, vrange, v):
    f = (np.asarray(v, dtype=F32) - F32(vmin)) / F32(vrange)
    f = np.clip(f, F32(0.0), F32(1.0))
    return (f * F32(65535.0) + F32(0.499)).astype(np.int64)


def _u16_to_float(vmin, vrange, u):
    return (F32(vmin) + F32(vrange) * F32(1.52590218966964e-05) * np.asarray(u).astype(F32)).astype(F32)


def compress_decompress(mat):
    """copy-feats --compress=true followed by a read: CompressedMatrix (kSpeechFeature, used for > 8 rows), float32.

    Global header: min, range.  Per column: the values at sorted positions 0, n/4, 3(n/4), n-1 as uint16, forced strictly
    increasing.  Per value: one byte, piecewise linear between those four points (64 / 128 / 63 steps)."""
    M = np.asarray(mat, dtype=F32)
    T, D = M.shape
    if T <= 8:
        raise ValueError("the speech-feature format is used for matrices with more than 8 rows")
    vmin, vmax = F32(M.min()), F32(M.max())
    if vmax == vmin:
        vmax = F32(vmin + F32(1.0) + abs(vmin))
    vrange = F32(vmax - vmin)
    out = np.empty_like(M)
    q = T // 4
    for d in range(D):
        col = M[:, d]
        s = np.sort(col)
        p0 = min(int(_float_to_u16(vmin, vrange, s[0])), 65532)
        p25 = min(max(int(_float_to_u16(vmin, vrange, s[q])), p0 + 1), 65533)
        p75 = min(max(int(_float_to_u16(vmin, vrange, s[3 * q])), p25 + 1), 65534)
        p100 = max(int(_float_to_u16(vmin, vrange, s[T - 1])), p75 + 1)
        f0, f25, f75, f100 = (_u16_to_float(vmin, vrange, p) for p in (p0, p25, p75, p100))
        c = np.empty(T, dtype=np.int64)
        lo = col < f25
        mid = (~lo) & (col < f75)
        hi = ~(lo | mid)
        with np.errstate(invalid="ignore", divide="ignore"):
            c[lo] = np.clip(((col[lo] - f0) / (f25 - f0) * F32(64.0) + F32(0.5)).astype(np.int64), 0, 64)
            c[mid] = 64 + np.clip(((col[mid] - f25) / (f75 - f25) * F32(128.0) + F32(0.5)).astype(np.int64), 0, 128)
            c[hi] = 192 + np.clip(((col[hi] - f75) / (f100 - f75) * F32(63.0) + F32(0.5)).astype(np.int64), 0, 63)
        cf = c.astype(F32)
        dec = np.where(c <= 64, f0 + (f25 - f0) * cf * F32(1.0 / 64.0),
                       np.where(c <= 192, f25 + (f75 - f25) * (cf - F32(64.0)) * F32(1.0 / 128.0),
                                f75 + (f100 - f75) * (cf - F32(192.0)) * F32(1.0 / 63.0)))
        out[:, d] = dec.astype(F32)
    return out


# ------------------------------------------------------------------------------------------------ libc generators
class GlibcRand:
    """glibc rand() with the default state (TYPE_3 additive feedback, degree 31, separation 3), i.e. srand(seed)."""

    def __init__(self, seed=1):
        r = [0] * 34
        r[0] = seed if seed != 0 else 1
        for i in range(1, 31):
            # 16807 * r[i-1] % 2147483647 computed without overflow (Schrage), as glibc's srandom_r does
            hi, lo = divmod(r[i - 1], 127773)
            w = 16807 * lo - 2836 * hi
            r[i] = w + 2147483647 if w < 0 else w
        for i in range(31, 34):
            r.append(0)
        self.r = [x & 0xFFFFFFFF for x in r[:31]]
        self.f, self.b = 3, 0                     # front / rear pointers
        for _ in range(310):
            self._step()

    def _step(self):
        self.r[self.f] = (self.r[self.f] + self.r[self.b]) & 0xFFFFFFFF
        out = self.r[self.f] >> 1
        self.f = (self.f + 1) % 31
        self.b = (self.b + 1) % 31
        return out

    def rand(self):
        return self._step()


def rand_r(seed):
    """glibc rand_r: -> (value in [0, RAND_MAX], new seed)."""
    nxt = seed & 0xFFFFFFFF
    nxt = (nxt * 1103515245 + 12345) & 0xFFFFFFFF
    result = (nxt // 65536) % 2048
    nxt = (nxt * 1103515245 + 12345) & 0xFFFFFFFF
    result = (result << 10) ^ ((nxt // 65536) % 1024)
    nxt = (nxt * 1103515245 + 12345) & 0xFFFFFFFF
    result = (result << 10) ^ ((nxt // 65536) % 1024)
    return result, nxt


def rand_r_stream(seed, n):
    """n successive rand_r values from `seed` (vectorised: the state is a plain LCG, three steps per value)."""
    a, c, m = 1103515245, 12345, 1 << 32
    k = np.arange(1, 3 * n + 1, dtype=object)
    # x_k = a^k x_0 + c (a^k - 1) / (a - 1)  (mod 2^32), evaluated iteratively in uint64 to stay exact
    x = np.empty(3 * n, dtype=np.uint64)
    cur = np.uint64(seed & 0xFFFFFFFF)
    A, Cc, M = np.uint64(a), np.uint64(c), np.uint64(0xFFFFFFFF)
    for i in range(3 * n):
        cur = (cur * A + Cc) & M
        x[i] = cur
    del k
    x = x.reshape(n, 3)
    hi = (x // np.uint64(65536))
    res = ((hi[:, 0] % np.uint64(2048)) << np.uint64(20)) ^ ((hi[:, 1] % np.uint64(1024)) << np.uint64(10)) ^ (hi[:, 2] % np.uint64(1024))
    return res.astype(np.int64)


def dither_frames(frames, dither=1.0, process_rand=None):
    """feature-window.cc Dither() applied to every extracted frame in order, like one compute-mfcc-feats process does:
    RandomState rstate (seed = rand() + 27437, one process-wide rand() per frame); data[i] += RandGauss(&rstate) * dither with
    RandGauss = sqrtf(-2 log(RandUniform)) * cosf(2 pi RandUniform), RandUniform = (rand_r + 1) / (RAND_MAX + 2), the uniform
    under the logarithm drawn first (operand order of the product as compiled by gcc)."""
    fr = np.asarray(frames, dtype=F32).copy()
    T, L = fr.shape
    g = process_rand or GlibcRand(1)
    for t in range(T):
        seed = (g.rand() + 27437) & 0xFFFFFFFF
        u = (rand_r_stream(seed, 2 * L).astype(np.float64) + 1.0) / (RAND_MAX + 2.0)
        u1, u2 = u[0::2], u[1::2]
        gauss = (np.sqrt((-2.0 * np.log(u1)).astype(F32)).astype(F32) * np.cos((2.0 * np.pi * u2).astype(F32)).astype(F32)).astype(F32)
        fr[t] = (fr[t] + gauss * F32(dither)).astype(F32)
    return fr
