"""Feature-pipeline configuration for the B200 scoring path.

Mirrors what the reference passes to Kaldi: ``pre-models/conf/mfcc.conf``
(``gmm_ubm_kaldiHelper.py:138``), ``pre-models/conf/vad.conf`` (``:158``) and
``pre-models/delta_opts`` (``:191-193``).  Files are parsed when present; otherwise
the egs/voxceleb/v1 values are used.
"""
import os
from dataclasses import dataclass, asdict

from . import kaldi_io


@dataclass
class FeatureConfig:
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
    dither: float = 0.0            # Kaldi's default 1.0 is a documented non-ideal effect; off here
    use_energy: bool = True        # Kaldi defaults below; the kernels implement exactly these values
    raw_energy: bool = True
    energy_floor: float = 0.0
    window_type: str = "povey"
    remove_dc_offset: bool = True
    htk_compat: bool = False
    round_to_power_of_two: bool = True
    vad_energy_threshold: float = 5.5
    vad_energy_mean_scale: float = 0.5
    vad_proportion_threshold: float = 0.12
    vad_frames_context: int = 2
    delta_window: int = 3
    delta_order: int = 2
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

    @property
    def feat_dim(self):
        return self.num_ceps * (self.delta_order + 1)

    def num_frames(self, n_samples):
        if self.snip_edges:
            if n_samples < self.frame_length:
                return 0
            return 1 + (n_samples - self.frame_length) // self.frame_shift
        return (n_samples + self.frame_shift // 2) // self.frame_shift

    def as_dict(self):
        return asdict(self)

    def check_supported(self):
        """The sm_100a kernels are specialised for the voxceleb/v1 recipe shapes."""
        if self.frame_length != 400 or self.frame_shift != 160 or self.padded_length != 512:
            raise ValueError("unsupported framing: the CUDA MFCC kernel is built for 25 ms / 10 ms at 16 kHz")
        if self.num_ceps != 24 or self.num_mel_bins > 32 or self.delta_order != 2 or self.delta_window != 3:
            raise ValueError("unsupported MFCC/delta shape: need 24 ceps, <=32 mel bins, delta order 2 window 3")
        if self.snip_edges:
            raise ValueError("snip_edges=true is not supported by the CUDA MFCC kernel")
        if self.dither != 0.0:
            raise ValueError("dither is not supported on the device path (SURVEY.md A.9)")
        fixed = {"use_energy": True, "raw_energy": True, "energy_floor": 0.0, "window_type": "povey", "remove_dc_offset": True,
                 "htk_compat": False, "round_to_power_of_two": True}
        for k, want in fixed.items():
            if getattr(self, k) != want:
                raise ValueError("unsupported mfcc.conf option: --%s=%s (the CUDA MFCC kernel implements %s only)"
                                 % (k.replace("_", "-"), getattr(self, k), want))


_BOOL = {"true": True, "false": False}


def load_feature_config(pre_model_dir):
    cfg = FeatureConfig()
    mfcc_conf = os.path.join(pre_model_dir, "conf", "mfcc.conf")
    if os.path.exists(mfcc_conf):
        o = kaldi_io.parse_conf(mfcc_conf)
        cfg.sample_frequency = float(o.get("sample-frequency", cfg.sample_frequency))
        cfg.frame_length_ms = float(o.get("frame-length", cfg.frame_length_ms))
        cfg.frame_shift_ms = float(o.get("frame-shift", cfg.frame_shift_ms))
        cfg.low_freq = float(o.get("low-freq", cfg.low_freq))
        cfg.high_freq = float(o.get("high-freq", cfg.high_freq))
        cfg.num_mel_bins = int(o.get("num-mel-bins", cfg.num_mel_bins))
        cfg.num_ceps = int(o.get("num-ceps", cfg.num_ceps))
        cfg.snip_edges = _BOOL[o.get("snip-edges", "true").lower()] if "snip-edges" in o else cfg.snip_edges
        cfg.preemph = float(o.get("preemphasis-coefficient", cfg.preemph))
        cfg.cepstral_lifter = float(o.get("cepstral-lifter", cfg.cepstral_lifter))
        cfg.dither = float(o.get("dither", cfg.dither))          # only an explicit value; Kaldi's implicit 1.0 stays off (A.9)
        cfg.energy_floor = float(o.get("energy-floor", cfg.energy_floor))
        cfg.window_type = o.get("window-type", cfg.window_type)
        for key, attr in (("use-energy", "use_energy"), ("raw-energy", "raw_energy"), ("remove-dc-offset", "remove_dc_offset"),
                          ("htk-compat", "htk_compat"), ("round-to-power-of-two", "round_to_power_of_two")):
            if key in o:
                setattr(cfg, attr, _BOOL[o[key].lower()])
        known = {"sample-frequency", "frame-length", "frame-shift", "low-freq", "high-freq", "num-mel-bins", "num-ceps", "snip-edges",
                 "preemphasis-coefficient", "cepstral-lifter", "dither", "energy-floor", "window-type", "use-energy", "raw-energy",
                 "remove-dc-offset", "htk-compat", "round-to-power-of-two"}
        unknown = sorted(set(o) - known)
        if unknown:
            import warnings
            warnings.warn("mfcc.conf options not understood by fakebob_b200 (ignored): %s" % ", ".join(unknown))
    vad_conf = os.path.join(pre_model_dir, "conf", "vad.conf")
    if os.path.exists(vad_conf):
        o = kaldi_io.parse_conf(vad_conf)
        cfg.vad_energy_threshold = float(o.get("vad-energy-threshold", cfg.vad_energy_threshold))
        cfg.vad_energy_mean_scale = float(o.get("vad-energy-mean-scale", cfg.vad_energy_mean_scale))
        cfg.vad_proportion_threshold = float(o.get("vad-proportion-threshold", cfg.vad_proportion_threshold))
        cfg.vad_frames_context = int(o.get("vad-frames-context", cfg.vad_frames_context))
    delta_opts = os.path.join(pre_model_dir, "delta_opts")
    if os.path.exists(delta_opts):
        with open(delta_opts) as f:
            txt = f.read().strip()
        for tok in txt.split():
            k, _, v = tok.lstrip("-").partition("=")
            if k == "delta-window":
                cfg.delta_window = int(v)
            elif k == "delta-order":
                cfg.delta_order = int(v)
    return cfg
