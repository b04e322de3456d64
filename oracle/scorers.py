"""Oracle (TEST INFRASTRUCTURE): the six task wrappers' score()/make_decisions() rules on top of
the oracle Kaldi arithmetic.  PARITY UNPINNED for the Kaldi part (see oracle/__init__.py).

Shape / normalisation rules follow the reference wrappers:
  gmm_OSI  gmm_ubm_OSI.py:50-112   score = LL_spk - LL_ubm, (B,K) or (K,) when B == 1; reject below threshold
  gmm_SV   gmm_ubm_SV.py:38-92     score = LL_spk - LL_ubm, (B,) or scalar
  gmm_CSI  gmm_ubm_CSI.py:54-110   score = (LL_spk - z_mean) / z_std, argmax
  iv_*     ivector_PLDA_{OSI,CSI,SV}.py + ivector_PLDA_kaldiHelper.py:282-308 (resolve_score squeeze rules)
"""
import numpy as np

from . import kaldi_feats as kf


def to_audio_list(audios, bits_per_sample=16):
    """gmm_ubm_OSI.py:70-85 input conventions -> list of int16 1-D arrays."""
    if isinstance(audios, np.ndarray):
        if audios.ndim == 1 or (audios.ndim == 2 and (audios.shape[0] == 1 or audios.shape[1] == 1)):
            lst = [audios.reshape(-1)]
        elif audios.ndim == 2:
            lst = [audios[:, i] for i in range(audios.shape[1])]
        else:
            raise ValueError("audios must be 1-D or 2-D")
    else:
        lst = [np.asarray(a).reshape(-1) for a in audios]
    return [kf.float_to_int16(a, bits_per_sample) for a in lst]


class _GmmBase:
    def __init__(self, gmms, cfg=None):
        self.gmms = gmms                      # list of oracle.diag_gmm.DiagGmm
        self.cfg = cfg or kf.FeatConfig()

    def raw_scores(self, audios, bits_per_sample=16):
        """(B, n_models) float64 average frame log-likelihoods (gmm_ubm_kaldiHelper.py:236-248)."""
        lst = to_audio_list(audios, bits_per_sample)
        out = np.zeros((len(lst), len(self.gmms)), dtype=np.float64)
        for b, w in enumerate(lst):
            X = kf.voiced_features(w, self.cfg)
            for k, g in enumerate(self.gmms):
                out[b, k] = g.avg_loglike(X)
        return out


class OracleGmmOSI(_GmmBase):
    def __init__(self, ubm, spk_gmms, threshold=0.0, cfg=None):
        super().__init__([ubm] + list(spk_gmms), cfg)
        self.threshold = threshold
        self.n_speakers = len(spk_gmms)

    def score(self, audios, fs=16000, bits_per_sample=16, debug=False, n_jobs=5):
        s = self.raw_scores(audios, bits_per_sample)
        final = s[:, 1:] - s[:, 0:1]
        return final if final.shape[0] > 1 else final[0]

    def make_decisions(self, audios, fs=16000, bits_per_sample=16, n_jobs=5, debug=False):
        score = self.score(audios, bits_per_sample=bits_per_sample)
        if score.ndim == 1:
            score = score[np.newaxis, :]
        decisions = list(np.argmax(score, axis=1))
        for i, v in enumerate(np.max(score, axis=1)):
            if v < self.threshold:
                decisions[i] = -1
        if score.shape[0] == 1:
            return decisions[0], score.flatten()
        return decisions, score


class OracleGmmSV(_GmmBase):
    def __init__(self, ubm, spk_gmm, threshold=0.0, cfg=None):
        super().__init__([ubm, spk_gmm], cfg)
        self.threshold = threshold

    def score(self, audios, fs=16000, bits_per_sample=16, debug=False, n_jobs=5):
        s = self.raw_scores(audios, bits_per_sample)
        final = s[:, 1] - s[:, 0]
        return final if final.shape[0] > 1 else final[0]

    def make_decisions(self, audios, fs=16000, bits_per_sample=16, n_jobs=5, debug=False):
        score = self.score(audios, bits_per_sample=bits_per_sample)
        if isinstance(score, np.ndarray):
            return [1 if v >= self.threshold else -1 for v in score], score
        return (1 if score >= self.threshold else -1), score


class OracleGmmCSI(_GmmBase):
    def __init__(self, spk_gmms, z_means, z_stds, cfg=None):
        super().__init__(list(spk_gmms), cfg)
        self.z_norm_means = np.asarray(z_means, dtype=np.float64)
        self.z_norm_stds = np.asarray(z_stds, dtype=np.float64)
        self.n_speakers = len(spk_gmms)

    def score(self, audios, fs=16000, bits_per_sample=16, debug=False, n_jobs=5):
        s = self.raw_scores(audios, bits_per_sample)
        final = (s - self.z_norm_means) / self.z_norm_stds
        return final if final.shape[0] > 1 else final[0]

    def make_decisions(self, audios, fs=16000, bits_per_sample=16, n_jobs=5, debug=False):
        score = self.score(audios, bits_per_sample=bits_per_sample)
        if score.ndim == 1:
            score = score[np.newaxis, :]
        decisions = list(np.argmax(score, axis=1))
        if score.shape[0] == 1:
            return decisions[0], score.flatten()
        return decisions, score


class _IvBase:
    def __init__(self, system, enrolled, z_means, z_stds):
        self.system = system                  # oracle.ivector.IvectorSystem
        self.enrolled = np.atleast_2d(np.asarray(enrolled, dtype=np.float64))
        self.z_norm_means = np.asarray(z_means, dtype=np.float64)
        self.z_norm_stds = np.asarray(z_stds, dtype=np.float64)

    def raw_scores(self, audios, bits_per_sample=16):
        lst = to_audio_list(audios, bits_per_sample)
        ivs = np.stack([self.system.extract(w) for w in lst])
        return self.system.plda_scores(self.enrolled, ivs)     # (B, K) float64


class OracleIvOSI(_IvBase):
    def __init__(self, system, enrolled, z_means, z_stds, threshold=0.0):
        super().__init__(system, enrolled, z_means, z_stds)
        self.threshold = threshold
        self.n_speakers = self.enrolled.shape[0]

    def score(self, audios, fs=16000, bits_per_sample=16, n_jobs=10, debug=False):
        s = self.raw_scores(audios, bits_per_sample)
        # resolve_score (ivector_PLDA_kaldiHelper.py:300-301): 1-D when one speaker or one test
        if s.shape[0] == 1 or s.shape[1] == 1:
            s = s.reshape(-1)
        return (s - self.z_norm_means) / self.z_norm_stds

    def make_decisions(self, audios, fs=16000, bits_per_sample=16, n_jobs=10, debug=False):
        s = self.score(audios, bits_per_sample=bits_per_sample)
        if s.ndim == 1:
            s = s[np.newaxis, :]
        dec = np.argmax(s, axis=1)
        for i, v in enumerate(np.max(s, axis=1)):
            if v < self.threshold:
                dec[i] = -1
        dec = list(dec)
        if len(dec) == 1:
            return dec[0], s.flatten()
        return dec, s


class OracleIvCSI(OracleIvOSI):
    def __init__(self, system, enrolled, z_means, z_stds):
        super().__init__(system, enrolled, z_means, z_stds, threshold=-np.inf)

    def make_decisions(self, audios, fs=16000, bits_per_sample=16, n_jobs=10, debug=False):
        s = self.score(audios, bits_per_sample=bits_per_sample)
        if s.ndim == 1:
            s = s[np.newaxis, :]
        dec = list(np.argmax(s, axis=1))
        if len(dec) == 1:
            return dec[0], s.flatten()
        return dec, s


class OracleIvSV(_IvBase):
    def __init__(self, system, enrolled, z_mean, z_std, threshold=0.0):
        super().__init__(system, enrolled, [z_mean], [z_std])
        self.threshold = threshold

    def score(self, audios, fs=16000, bits_per_sample=16, n_jobs=10, debug=False):
        s = self.raw_scores(audios, bits_per_sample).reshape(-1)
        s = (s - self.z_norm_means[0]) / self.z_norm_stds[0]
        return s if s.size > 1 else s[0]

    def make_decisions(self, audios, fs=16000, bits_per_sample=16, n_jobs=10, debug=False):
        s = self.score(audios, bits_per_sample=bits_per_sample)
        if isinstance(s, np.ndarray):
            return [1 if v >= self.threshold else -1 for v in s], s
        return (1 if s >= self.threshold else -1), s
