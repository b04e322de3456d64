"""Oracle (TEST INFRASTRUCTURE): numpy restatement of the Kaldi feature pipeline the
reference shells out to.  PARITY UNPINNED against real Kaldi (see oracle/__init__.py).

Reference call sites:
  compute-mfcc-feats ........ gmm_ubm_kaldiHelper.py:138  (steps/make_mfcc.sh, conf/mfcc.conf)
  compute-vad ............... gmm_ubm_kaldiHelper.py:158  (sid/compute_vad_decision.sh, conf/vad.conf)
  add-deltas | apply-cmvn-sliding --norm-vars=false --center=true --cmn-window=300
             | select-voiced-frames ... gmm_ubm_kaldiHelper.py:195-198
Upstream algorithm: SURVEY.md Appendix A.1-A.6 (feat/feature-window.cc,
feat/feature-mfcc.cc, feat/mel-computations.cc, ivector/voice-activity-detection.cc,
feat/feature-functions.cc).

Arithmetic is float32 where Kaldi's BaseFloat is float, float64 where Kaldi
accumulates in double (Vector::Sum, sliding-CMN running sums).
"""
from dataclasses import dataclass

import numpy as np

F32 = np.float32
FLT_EPS = np.finfo(np.float32).eps


@dataclass
class FeatConfig:
    # conf/mfcc.conf of egs/voxceleb/v1 + Kaldi defaults
    sample_frequency: float = 16000.0
    frame_length_ms: float = 25.0
    frame_shift_ms: float = 10.0
    low_freq: float = 20.0
    high_freq: float = 7600.0
    num_mel_bins: int = 30
    num_ceps: int = 24
    snip_edges: bool = False
    preemph: float = 0.97
    cepstral_lifter: float = 22.0
    dither: float = 0.0          # Kaldi default is 1.0 (A.9); OFF for GPU parity
    # conf/vad.conf
    vad_energy_threshold: float = 5.5
    vad_energy_mean_scale: float = 0.5
    vad_proportion_threshold: float = 0.12
    vad_frames_context: int = 2
    # delta_opts
    delta_window: int = 3
    delta_order: int = 2
    # apply-cmvn-sliding
    cmn_window: int = 300

    @property
    def frame_length(self):
        return int(self.sample_frequency * 0.001 * self.frame_length_ms)

    @property
    def frame_shift(self):
        return int(self.sample_frequency * 0.001 * self.frame_shift_ms)

    @property
    def padded_length(self):
        n = 1
        while n < self.frame_length:
            n *= 2
        return n


def num_frames(n_samples, cfg):
    """feature-window.cc NumFrames()."""
    if cfg.snip_edges:
        if n_samples < cfg.frame_length:
            return 0
        return 1 + (n_samples - cfg.frame_length) // cfg.frame_shift
    return (n_samples + cfg.frame_shift // 2) // cfg.frame_shift


def povey_window(cfg):
    n = cfg.frame_length
    a = 2.0 * np.pi / (n - 1)
    i = np.arange(n, dtype=np.float64)
    return np.power(0.5 - 0.5 * np.cos(a * i), 0.85).astype(F32)


def mel_scale(f):
    return F32(1127.0) * np.log(F32(1.0) + np.asarray(f, dtype=F32) / F32(700.0)).astype(F32)


def mel_banks(cfg):
    """mel-computations.cc MelBanks::MelBanks (no VTLN). Returns dense (num_bins, nfft/2) float32."""
    nfft = cfg.padded_length
    n_fft_bins = nfft // 2
    fft_bin_width = F32(cfg.sample_frequency / nfft)
    mel_low = mel_scale(cfg.low_freq)
    mel_high = mel_scale(cfg.high_freq)
    delta = F32((mel_high - mel_low) / F32(cfg.num_mel_bins + 1))
    W = np.zeros((cfg.num_mel_bins, n_fft_bins), dtype=F32)
    mel = mel_scale(fft_bin_width * np.arange(n_fft_bins, dtype=F32))
    for b in range(cfg.num_mel_bins):
        left = F32(mel_low + F32(b) * delta)
        center = F32(mel_low + F32(b + 1) * delta)
        right = F32(mel_low + F32(b + 2) * delta)
        for i in range(n_fft_bins):
            m = mel[i]
            if m > left and m < right:
                if m <= center:
                    W[b, i] = (m - left) / (center - left)
                else:
                    W[b, i] = (right - m) / (right - center)
    return W


def dct_matrix(cfg):
    """matrix-functions.cc ComputeDctMatrix, first num_ceps rows."""
    N = cfg.num_mel_bins
    M = np.zeros((N, N), dtype=np.float64)
    M[0, :] = np.sqrt(1.0 / N)
    n = np.arange(N, dtype=np.float64)
    for k in range(1, N):
        M[k, :] = np.sqrt(2.0 / N) * np.cos(np.pi / N * (n + 0.5) * k)
    return M[: cfg.num_ceps].astype(F32)


def lifter_coeffs(cfg):
    Q = cfg.cepstral_lifter
    i = np.arange(cfg.num_ceps, dtype=np.float64)
    return (1.0 + 0.5 * Q * np.sin(np.pi * i / Q)).astype(F32)


def extract_frames(wave, cfg):
    """feature-window.cc ExtractWindow with snip_edges=false reflection. wave: int16/float (N,) -> (T, frame_length) f32."""
    wave = np.asarray(wave).astype(F32).reshape(-1)
    N = wave.shape[0]
    T = num_frames(N, cfg)
    L, S = cfg.frame_length, cfg.frame_shift
    if cfg.snip_edges:
        start = S * np.arange(T)
    else:
        start = S * np.arange(T) + S // 2 - L // 2
    idx = start[:, None] + np.arange(L)[None, :]
    # reflect until in range (Kaldi's while loop)
    for _ in range(8):
        neg = idx < 0
        idx = np.where(neg, -idx - 1, idx)
        big = idx >= N
        idx = np.where(big, 2 * N - 1 - idx, idx)
        if not (neg.any() or big.any()):
            break
    return wave[idx]


def mfcc(wave, cfg=None, rng=None, frames=None):
    """compute-mfcc-feats: int16-valued samples -> (T, num_ceps) float32 (use_energy, raw_energy).
    frames: already extracted (T, L) windows, e.g. with Kaldi's libc dither applied (oracle.kaldi_nonideal.dither_frames)."""
    cfg = cfg or FeatConfig()
    fr = extract_frames(wave, cfg) if frames is None else np.asarray(frames, dtype=F32)   # (T, L) f32
    T, L = fr.shape
    if T == 0:
        return np.zeros((0, cfg.num_ceps), dtype=F32)
    if cfg.dither != 0.0 and frames is None:
        rng = rng or np.random.default_rng(0)
        fr = (fr + F32(cfg.dither) * rng.standard_normal(fr.shape).astype(F32)).astype(F32)
    # remove_dc_offset: Sum() in double, cast to float, / L in float
    mean = (fr.astype(np.float64).sum(axis=1).astype(F32) / F32(L)).astype(F32)
    fr = (fr - mean[:, None]).astype(F32)
    # raw log energy (before pre-emphasis and windowing)
    energy = np.maximum(np.einsum("tl,tl->t", fr, fr, dtype=F32), FLT_EPS).astype(F32)
    log_energy = np.log(energy).astype(F32)
    # pre-emphasis
    pe = F32(cfg.preemph)
    out = np.empty_like(fr)
    out[:, 1:] = fr[:, 1:] - pe * fr[:, :-1]
    out[:, 0] = fr[:, 0] - pe * fr[:, 0]
    fr = (out * povey_window(cfg)[None, :]).astype(F32)
    nfft = cfg.padded_length
    spec = np.fft.rfft(fr, n=nfft, axis=1)               # complex64 for f32 input (numpy>=2)
    power = (spec.real.astype(F32) ** 2 + spec.imag.astype(F32) ** 2).astype(F32)   # (T, nfft/2+1)
    W = mel_banks(cfg)
    mel_e = (power[:, : nfft // 2] @ W.T).astype(F32)
    mel_e = np.log(np.maximum(mel_e, FLT_EPS)).astype(F32)
    feat = (mel_e @ dct_matrix(cfg).T).astype(F32)
    feat = (feat * lifter_coeffs(cfg)[None, :]).astype(F32)
    feat[:, 0] = log_energy
    return feat


def compute_vad(feats, cfg=None):
    """voice-activity-detection.cc ComputeVadEnergy on column 0. -> (T,) float32 of {0,1}."""
    cfg = cfg or FeatConfig()
    T = feats.shape[0]
    if T == 0:
        return np.zeros((0,), dtype=F32)
    le = feats[:, 0].astype(F32)
    thr = F32(cfg.vad_energy_threshold)
    if cfg.vad_energy_mean_scale != 0.0:
        s = F32(le.astype(np.float64).sum())
        thr = F32(thr + F32(cfg.vad_energy_mean_scale) * s / F32(T))
    above = (le > thr).astype(np.int64)
    c = cfg.vad_frames_context
    pad = np.concatenate([np.zeros(c, np.int64), above, np.zeros(c, np.int64)])
    ones = np.concatenate([np.zeros(c, np.int64), np.ones(T, np.int64), np.zeros(c, np.int64)])
    k = np.ones(2 * c + 1, dtype=np.int64)
    num = np.convolve(pad, k, mode="valid")
    den = np.convolve(ones, k, mode="valid")
    return (num.astype(F32) >= den.astype(F32) * F32(cfg.vad_proportion_threshold)).astype(F32)


def delta_scales(cfg):
    """feature-functions.cc DeltaFeatures::DeltaFeatures."""
    scales = [np.array([1.0], dtype=F32)]
    w = cfg.delta_window
    for _ in range(cfg.delta_order):
        prev = scales[-1]
        prev_off = (len(prev) - 1) // 2
        cur_off = prev_off + w
        cur = np.zeros(len(prev) + 2 * w, dtype=F32)
        normalizer = F32(0.0)
        for j in range(-w, w + 1):
            normalizer += F32(j * j)
            for k in range(-prev_off, prev_off + 1):
                cur[j + k + cur_off] += F32(j) * prev[k + prev_off]
        cur = (cur * (F32(1.0) / normalizer)).astype(F32)
        scales.append(cur)
    return scales


def add_deltas(feats, cfg=None):
    """add-deltas: (T, d) -> (T, d*(order+1)) float32; frame index clamped on the static features."""
    cfg = cfg or FeatConfig()
    T, d = feats.shape
    out = np.zeros((T, d * (cfg.delta_order + 1)), dtype=F32)
    t = np.arange(T)
    for i, sc in enumerate(delta_scales(cfg)):
        mo = (len(sc) - 1) // 2
        acc = np.zeros((T, d), dtype=F32)
        for j in range(-mo, mo + 1):
            s = sc[j + mo]
            if s != 0.0:
                acc = (acc + s * feats[np.clip(t + j, 0, T - 1)]).astype(F32)
        out[:, i * d:(i + 1) * d] = acc
    return out


def sliding_cmn(feats, cfg=None):
    """apply-cmvn-sliding --norm-vars=false --center=true --cmn-window=W (feature-functions.cc SlidingWindowCmn)."""
    cfg = cfg or FeatConfig()
    T, d = feats.shape
    out = np.empty_like(feats, dtype=F32)
    if T == 0:
        return out
    P = np.concatenate([np.zeros((1, d)), np.cumsum(feats.astype(np.float64), axis=0)], axis=0)
    W = cfg.cmn_window
    t = np.arange(T)
    ws = t - W // 2
    we = ws + W
    shift = np.where(ws < 0, -ws, 0)
    ws = ws + shift
    we = we + shift
    over = np.where(we > T, we - T, 0)
    ws = np.maximum(ws - over, 0)
    we = we - over
    n = (we - ws)
    cur_sum = P[we] - P[ws]                                  # double window sums
    alpha = (F32(-1.0) / n.astype(F32)).astype(F32).astype(np.float64)   # BaseFloat alpha
    out[:] = (feats.astype(np.float64) + alpha[:, None] * cur_sum).astype(F32)
    return out


def select_voiced(feats, vad):
    return feats[vad != 0]


def voiced_features(wave, cfg=None):
    """Whole front-end: int16 samples -> (Tv, 72) float32 as seen by gmm-global-get-frame-likes."""
    cfg = cfg or FeatConfig()
    m = mfcc(wave, cfg)
    v = compute_vad(m, cfg)
    f = sliding_cmn(add_deltas(m, cfg), cfg)
    return select_voiced(f, v)


def float_to_int16(audio, bits_per_sample=16):
    """gmm_ubm_OSI.py:83-85 : (audio * 2**(bits-1)).astype(np.int16) -- truncation toward zero."""
    audio = np.asarray(audio)
    if audio.dtype == np.int16:
        return audio
    return (audio * (2 ** (bits_per_sample - 1))).astype(np.int16)
