"""i-vector / PLDA task wrappers with the reference's Python API, scored on the B200.

Same class names, constructor signatures, speaker ordering (sorted by spk_id string, ivector_PLDA_OSI.py:56-57),
z-normalisation ((score - z_mean) / z_std, ivector_PLDA_OSI.py:119) and output squeeze rules
(resolve_score, ivector_PLDA_kaldiHelper.py:300-301: 1-D when one speaker or one test) as
  iv_OSI  /root/reference ivector_PLDA_OSI.py:16-143
  iv_CSI  ivector_PLDA_CSI.py:18-135
  iv_SV   ivector_PLDA_SV.py:20-119
but ``score()`` makes one C-ABI call (``fb_score_ivector_host``) instead of driving ``ivector_PLDA_kaldiHelper.score``.
"""
import os

import numpy as np

from . import kaldi_io
from .config import load_feature_config
from .engine import IvectorEngine, to_audio_list


class _IvScorerBase(object):
    _fb_arch = "iv"

    def _setup_engine(self, identity_locations, pre_model_dir, device=None):
        self.pre_model_dir = os.path.abspath(pre_model_dir)
        self.feat_cfg = load_feature_config(self.pre_model_dir)
        self._engine = IvectorEngine(self.pre_model_dir, feat_cfg=self.feat_cfg, device=device)
        enrolled = np.stack([np.asarray(kaldi_io.read_vector(loc), dtype=np.float32) for loc in identity_locations])
        self._engine.set_enrolled(enrolled)

    def _plda(self, audios, bits_per_sample):
        return self._engine.score_plda(to_audio_list(audios, bits_per_sample))

    @staticmethod
    def _order(spk_ids, *cols):
        idx = sorted(range(len(spk_ids)), key=lambda i: spk_ids[i])
        return [[c[i] for i in idx] for c in (spk_ids,) + cols]

    def _parse(self, model_list):
        spk = [m[0] for m in model_list]
        utt = [m[1] for m in model_list]
        loc = [m[2] for m in model_list]
        zm = [m[3] for m in model_list]
        zs = [m[4] for m in model_list]
        self.spk_ids, self.utt_ids, self.identity_locations, zm, zs = self._order(spk, utt, loc, zm, zs)
        self.z_norm_means = np.array(zm, dtype=np.float64)
        self.z_norm_stds = np.array(zs, dtype=np.float64)
        self.n_speakers = len(model_list)

    def _write_scp(self, base):
        self.train_ivector_scp = base + "/ivector.scp"
        with open(self.train_ivector_scp, "w") as f:
            for u, l in zip(self.utt_ids, self.identity_locations):
                f.write("%s %s\n" % (u, l))


class iv_OSI(_IvScorerBase):
    _fb_task = "OSI"

    def __init__(self, group_id, model_list, pre_model_dir="pre-models", threshold=0.0, device=None):
        self.group_id = os.path.abspath(group_id)
        os.makedirs(self.group_id, exist_ok=True)
        self.threshold = threshold
        self._parse(model_list)
        self._write_scp(self.group_id)
        self._setup_engine(self.identity_locations, pre_model_dir, device)

    def score(self, audio_list, fs=16000, bits_per_sample=16, n_jobs=10, debug=False):
        s = self._plda(audio_list, bits_per_sample)
        if s.shape[0] == 1 or s.shape[1] == 1:
            s = s.reshape(-1)
        return (s - self.z_norm_means) / self.z_norm_stds

    def make_decisions(self, audios, fs=16000, bits_per_sample=16, n_jobs=10, debug=False):
        reject = -1
        score_array = self.score(audios, fs=fs, bits_per_sample=bits_per_sample, n_jobs=n_jobs, debug=debug)
        if len(score_array.shape) == 1:
            score_array = score_array[np.newaxis, :]
        max_score = np.max(score_array, axis=1)
        decisions = np.argmax(score_array, axis=1)
        for i, s in enumerate(max_score):
            if s < self.threshold:
                decisions[i] = reject
        decisions = list(decisions)
        if len(decisions) == 1:
            decisions = decisions[0]
            score_array = score_array.flatten()
        return decisions, score_array


class iv_CSI(_IvScorerBase):
    _fb_task = "CSI"

    def __init__(self, group_id, model_list, pre_model_dir="pre-models", device=None):
        self.group_id = os.path.abspath(group_id)
        os.makedirs(self.group_id, exist_ok=True)
        self._parse(model_list)
        self._write_scp(self.group_id)
        self._setup_engine(self.identity_locations, pre_model_dir, device)

    def score(self, audio_list, fs=16000, bits_per_sample=16, n_jobs=10, debug=False):
        s = self._plda(audio_list, bits_per_sample)
        if s.shape[0] == 1 or s.shape[1] == 1:
            s = s.reshape(-1)
        return (s - self.z_norm_means) / self.z_norm_stds

    def make_decisions(self, audios, fs=16000, bits_per_sample=16, n_jobs=10, debug=False):
        score_array = self.score(audios, fs=fs, bits_per_sample=bits_per_sample, n_jobs=n_jobs, debug=debug)
        if len(score_array.shape) == 1:
            score_array = score_array[np.newaxis, :]
        decisions = list(np.argmax(score_array, axis=1))
        if len(decisions) == 1:
            decisions = decisions[0]
            score_array = score_array.flatten()
        return decisions, score_array


class iv_SV(_IvScorerBase):
    _fb_task = "SV"

    def __init__(self, spk_id, model, pre_model_dir="pre-models", threshold=0.0, device=None):
        self.spk_id = os.path.abspath(spk_id)
        os.makedirs(self.spk_id, exist_ok=True)
        self.threshold = threshold
        self.n_speakers = 1
        self.spk_ids = [model[0]]
        self.utt_id = model[1]
        self.utt_ids = [model[1]]
        self.identity_location = model[2]
        self.identity_locations = [model[2]]
        self.z_norm_mean = model[3]
        self.z_norm_std = model[4]
        self.z_norm_means = np.array([model[3]], dtype=np.float64)
        self.z_norm_stds = np.array([model[4]], dtype=np.float64)
        self._write_scp(self.spk_id)
        self._setup_engine([self.identity_location], pre_model_dir, device)

    def score(self, audio_list, fs=16000, bits_per_sample=16, n_jobs=10, debug=False):
        s = self._plda(audio_list, bits_per_sample).reshape(-1)
        s = (s - self.z_norm_mean) / self.z_norm_std
        return s if s.size > 1 else s[0]

    def make_decisions(self, audio_list, fs=16000, bits_per_sample=16, n_jobs=10, debug=False):
        accept, reject = 1, -1
        scores = self.score(audio_list, fs=fs, bits_per_sample=bits_per_sample, n_jobs=n_jobs, debug=debug)
        if isinstance(scores, np.ndarray):
            decisions = [accept if s >= self.threshold else reject for s in scores]
        else:
            decisions = accept if scores >= self.threshold else reject
        return decisions, scores

    def make_decisions_value(self, score):
        return -1 if score < self.threshold else 1
