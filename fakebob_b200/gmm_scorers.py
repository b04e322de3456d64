"""GMM-UBM task wrappers with the reference's Python API, scored on the B200.

Same class names, constructor signatures, ``score()`` / ``make_decisions()`` semantics, attributes
(``spk_ids``, ``utt_ids``, ``n_speakers``, ``threshold``, ``z_norm_means/stds``) and output squeeze rules as
  gmm_OSI  /root/reference gmm_ubm_OSI.py:13-112
  gmm_CSI  gmm_ubm_CSI.py:13-110
  gmm_SV   gmm_ubm_SV.py:13-92
but ``score()`` makes one C-ABI call (``fb_score_gmm_host``) instead of driving
``gmm_ubm_kaldiHelper.score`` (wav files + ~8 shell scripts + text parsing per call).
Errors raise (the reference ignores subprocess failures).
"""
import os

import numpy as np

from . import kaldi_io
from .config import load_feature_config
from .engine import GmmEngine, to_audio_list


class _GmmScorerBase(object):
    _fb_arch = "gmm"

    def _setup_engine(self, model_paths, pre_model_dir, device=None):
        self.pre_model_dir = os.path.abspath(pre_model_dir)
        self.feat_cfg = load_feature_config(self.pre_model_dir)
        self.model_list = list(model_paths)
        self._engine = GmmEngine.from_files(self.model_list, feat_cfg=self.feat_cfg, device=device)

    def _avg_ll(self, audios, bits_per_sample):
        return self._engine.score_avg_ll(to_audio_list(audios, bits_per_sample))


class gmm_OSI(_GmmScorerBase):
    _fb_task = "OSI"

    def __init__(self, group_id, model_list, ubm, pre_model_dir="pre-models", threshold=0.0, device=None):
        self.group_id = os.path.abspath(group_id)
        if not os.path.exists(self.group_id):
            os.makedirs(self.group_id)
        self.threshold = threshold
        self.n_speakers = len(model_list)
        self.spk_ids = [m[0] for m in model_list]
        self.utt_ids = [m[1] for m in model_list]
        self.identity_locations = [m[2] for m in model_list]
        self._setup_engine([ubm] + self.identity_locations, pre_model_dir, device)

    def score(self, audios, fs=16000, bits_per_sample=16, debug=False, n_jobs=5):
        ll = self._avg_ll(audios, bits_per_sample)
        final_score = ll[:, 1:] - ll[:, 0:1]
        return final_score if final_score.shape[0] > 1 else final_score[0]

    def make_decisions(self, audios, fs=16000, bits_per_sample=16, n_jobs=5, debug=False):
        reject = -1
        score = self.score(audios, fs=fs, bits_per_sample=bits_per_sample, debug=debug, n_jobs=n_jobs)
        if len(score.shape) == 1:
            score = score[np.newaxis, :]
        max_score = np.max(score, axis=1)
        decisions = list(np.argmax(score, axis=1))
        for i, value in enumerate(max_score):
            if value < self.threshold:
                decisions[i] = reject
        if score.shape[0] == 1:
            decisions = decisions[0]
            score = score.flatten()
        return decisions, score


class gmm_SV(_GmmScorerBase):
    _fb_task = "SV"

    def __init__(self, spk_id, model, ubm, pre_model_dir="pre-models", threshold=0.0, device=None):
        self.spk_id = os.path.abspath(spk_id)
        if not os.path.exists(self.spk_id):
            os.makedirs(self.spk_id)
        self.threshold = threshold
        self.n_speakers = 1
        self.spk_ids = [model[0]]
        self.utt_id = model[1]
        self.identity_location = model[2]
        self._setup_engine([ubm, self.identity_location], pre_model_dir, device)

    def score(self, audios, fs=16000, bits_per_sample=16, debug=False, n_jobs=5):
        ll = self._avg_ll(audios, bits_per_sample)
        final_score = ll[:, 1] - ll[:, 0]
        return final_score if final_score.shape[0] > 1 else final_score[0]

    def make_decisions(self, audios, fs=16000, bits_per_sample=16, n_jobs=5, debug=False):
        accept, reject = 1, -1
        score = self.score(audios, fs=fs, bits_per_sample=bits_per_sample, debug=debug, n_jobs=n_jobs)
        if isinstance(score, np.ndarray):
            decisions = [accept if v >= self.threshold else reject for v in score]
        else:
            decisions = accept if score >= self.threshold else reject
        return decisions, score


class gmm_CSI(_GmmScorerBase):
    _fb_task = "CSI"

    def __init__(self, group_id, model_list, pre_model_dir="pre-models", device=None):
        self.group_id = os.path.abspath(group_id)
        if not os.path.exists(self.group_id):
            os.makedirs(self.group_id)
        self.n_speakers = len(model_list)
        self.spk_ids = [m[0] for m in model_list]
        self.utt_ids = [m[1] for m in model_list]
        self.identity_locations = [m[2] for m in model_list]
        self.z_norm_means = np.array([m[3] for m in model_list], dtype=np.float64)
        self.z_norm_stds = np.array([m[4] for m in model_list], dtype=np.float64)
        self._setup_engine(self.identity_locations, pre_model_dir, device)

    def score(self, audios, fs=16000, bits_per_sample=16, debug=False, n_jobs=5):
        ll = self._avg_ll(audios, bits_per_sample)
        final_score = (ll - self.z_norm_means) / self.z_norm_stds
        return final_score if final_score.shape[0] > 1 else final_score[0]

    def make_decisions(self, audios, fs=16000, bits_per_sample=16, n_jobs=5, debug=False):
        score = self.score(audios, fs=fs, bits_per_sample=bits_per_sample, debug=debug, n_jobs=n_jobs)
        if len(score.shape) == 1:
            score = score[np.newaxis, :]
        decisions = list(np.argmax(score, axis=1))
        if score.shape[0] == 1:
            decisions = decisions[0]
            score = score.flatten()
        return decisions, score
