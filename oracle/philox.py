"""Oracle (TEST INFRASTRUCTURE): Philox4x32-10 counter-based normals, the device RNG's CPU twin.

The reference draws NES noise with the global numpy generator
(``FAKEBOB.py:234``: ``np.random.normal(size=(N, samples_per_draw // 2))``), which costs
~92 ms per iteration on the host at S=50.  The timed device path
(``rng="philox"``) instead generates noise on the GPU; this file defines that stream
bit-for-bit on the integer side so the oracle can replay it:

  key      = (seed_lo, seed_hi)
  counter  = (n // 4, pair_index j, iter_lo, iter_hi)          one call -> 4 x uint32
  u_a      = ((x_{2q} >> 9) + 0.5) * 2^-23 ,  u_b = ((x_{2q+1} >> 9) + 0.5) * 2^-23     q = 0, 1   (exact in float32)
  z_{2q}   = sqrt(-2 ln u_a) * cos(2 pi u_b),  z_{2q+1} = sqrt(-2 ln u_a) * sin(2 pi u_b)     (float32 arithmetic)
  noise[n = 4*(n//4) + i, j] = float64(z_i)

The integer stream and the uniforms are exact; ``log``/``sin``/``cos`` are evaluated by each side's float32 libm
(<= 2 ulp apart, i.e. ~1e-7 relative), far below the int16 quantisation step the perturbed audio goes through
(``gmm_ubm_OSI.py:83-85``): a perturbed sample differs between the two sides with probability ~3e-6.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over counter arrays (uint32). Returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint32).copy()
    c1 = np.asarray(c1, dtype=np.uint32).copy()
    c2 = np.asarray(c2, dtype=np.uint32).copy()
    c3 = np.asarray(c3, dtype=np.uint32).copy()
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & MASK).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & MASK).astype(np.uint32)
            n0 = hi1 ^ c1 ^ k0
            n1 = lo1
            n2 = hi0 ^ c3 ^ k1
            n3 = lo0
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


F32 = np.float32
TWO_PI_F = F32(6.2831855)
SCALE_23 = F32(2.0 ** -23)


def normal_noise(seed, it, n_samples, n_pairs):
    """-> (n_samples, n_pairs) float64, the device ``rng='philox'`` stream for iteration ``it``.

    Box-Muller in float32 on 23-bit uniforms u = ((x >> 9) + 0.5) * 2^-23 (exact), widened to float64."""
    seed = int(seed)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    ng = (n_samples + 3) // 4
    g = np.arange(ng, dtype=np.uint32)[:, None]
    j = np.arange(n_pairs, dtype=np.uint32)[None, :]
    c0 = np.broadcast_to(g, (ng, n_pairs))
    c1 = np.broadcast_to(j, (ng, n_pairs))
    c2 = np.full((ng, n_pairs), int(it) & 0xFFFFFFFF, dtype=np.uint32)
    c3 = np.full((ng, n_pairs), (int(it) >> 32) & 0xFFFFFFFF, dtype=np.uint32)
    x0, x1, x2, x3 = philox4x32_10(c0, c1, c2, c3, k0, k1)
    out = np.empty((ng, 4, n_pairs), dtype=np.float64)
    for q, (xa, xb) in enumerate(((x0, x1), (x2, x3))):
        ua = (((xa >> np.uint32(9)).astype(F32) + F32(0.5)) * SCALE_23).astype(F32)
        ub = (((xb >> np.uint32(9)).astype(F32) + F32(0.5)) * SCALE_23).astype(F32)
        rad = np.sqrt((F32(-2.0) * np.log(ua)).astype(F32)).astype(F32)
        ang = (TWO_PI_F * ub).astype(F32)
        out[:, 2 * q, :] = (rad * np.cos(ang).astype(F32)).astype(F32)
        out[:, 2 * q + 1, :] = (rad * np.sin(ang).astype(F32)).astype(F32)
    return out.reshape(ng * 4, n_pairs)[:n_samples]
