"""Deterministic stand-in scorer with the reference's duck-typed interface (README.md:136), used to pin the
NES-loop oracle against the reference's own FAKEBOB.py without Kaldi.  Scores are a smooth function of the
int16-quantised audio (the quantisation mirrors gmm_ubm_OSI.py:83-85 so perturbations below 1 LSB vanish)."""
import numpy as np


class StubScorer:
    def __init__(self, n_speakers, n_samples, seed=3, sv=False, threshold=0.0):
        r = np.random.default_rng(seed)
        self.W = r.standard_normal((n_speakers, n_samples)) / np.sqrt(n_samples) * 4.0
        self.b = r.standard_normal(n_speakers) * 0.3
        self.sv = sv
        self.threshold = threshold
        self.n_speakers = n_speakers

    def score(self, audios, fs=16000, bits_per_sample=16, debug=False, n_jobs=5):
        a = np.asarray(audios)
        if a.ndim == 1:
            a = a[:, None]
        elif a.shape[0] == 1:
            a = a.T
        q = (a * 32768).astype(np.int16).astype(np.float64) / 32768.0
        s = (np.tanh(self.W @ q) * 3.0 + self.b[:, None]).T           # (B, K)
        if self.sv:
            s = s[:, 0]
            return s if s.shape[0] > 1 else s[0]
        return s if s.shape[0] > 1 else s[0]

    def make_decisions(self, audios, fs=16000, bits_per_sample=16, n_jobs=5, debug=False):
        s = self.score(audios)
        if self.sv:
            if isinstance(s, np.ndarray):
                return [1 if v >= self.threshold else -1 for v in s], s
            return (1 if s >= self.threshold else -1), s
        s2 = s[None, :] if s.ndim == 1 else s
        d = list(np.argmax(s2, axis=1))
        for i, v in enumerate(np.max(s2, axis=1)):
            if v < self.threshold:
                d[i] = -1
        if s2.shape[0] == 1:
            return d[0], s2.flatten()
        return d, s2


CASES = {
    # name: (task, attack_type, n_speakers, attack kwargs, FakeBob hyper-parameters)
    "osi_untargeted": ("OSI", "untargeted", 3, dict(threshold=2.56), dict(max_iter=40, samples_per_draw=10)),
    "osi_targeted": ("OSI", "targeted", 3, dict(threshold=2.6, target=2), dict(max_iter=40, samples_per_draw=10)),
    "csi_targeted": ("CSI", "targeted", 4, dict(target=0), dict(max_iter=25, samples_per_draw=12, adver_thresh=0.1)),
    "csi_untargeted": ("CSI", "untargeted", 4, dict(true=2), dict(max_iter=25, samples_per_draw=7)),
    "sv": ("SV", "untargeted", 1, dict(threshold=-0.78), dict(max_iter=40, samples_per_draw=10, plateau_length=3)),
}
N_SAMPLES = 4000


def make_audio(seed=0):
    r = np.random.default_rng(seed)
    a = r.uniform(-0.3, 0.3, N_SAMPLES)
    return np.trunc(a * 32768) / 32768
