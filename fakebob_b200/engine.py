"""Host-side engine: owns one ``fb_ctx`` (one GPU, one stream), the resident GMM parameters and the
NES attack state.  This is what replaces ``gmm_ubm_kaldiHelper`` on the hot path
(``gmm_ubm_kaldiHelper.py:270-291``): no wav/ark/text files, no subprocesses -- one C-ABI call per
batch of audios, or per K NES iterations.
"""
import ctypes as C
import os

import numpy as np

from . import _lib, kaldi_io
from .config import FeatureConfig


def default_device():
    for k in ("FAKEBOB_DEVICE", "LOCAL_RANK"):
        if k in os.environ:
            return int(os.environ[k])
    return 0


def to_audio_list(audios, bits_per_sample=16):
    """Input conventions of the reference's score() (gmm_ubm_OSI.py:70-85): ndarray (N,), (N,1), (1,N) is
    one audio; (N,B) is B audios as columns; otherwise a list of 1-D arrays of possibly different
    lengths.  Non-int16 data is scaled by 2**(bits-1) and truncated toward zero.  Single 2-D audios
    are flattened."""
    if isinstance(audios, np.ndarray):
        if audios.ndim == 1 or (audios.ndim == 2 and (audios.shape[0] == 1 or audios.shape[1] == 1)):
            lst = [audios.reshape(-1)]
        elif audios.ndim == 2:
            lst = [audios[:, i] for i in range(audios.shape[1])]
        else:
            raise ValueError("audios must be a 1-D or 2-D array or a list of 1-D arrays")
    else:
        lst = [np.asarray(a).reshape(-1) for a in audios]
    out = []
    for a in lst:
        if a.dtype != np.int16:
            a = (a * (2 ** (bits_per_sample - 1))).astype(np.int16)
        out.append(np.ascontiguousarray(a))
    return out


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class GmmEngine:
    """Resident diagonal GMMs + front-end on one B200."""

    def __init__(self, gmm_params, feat_cfg=None, device=None, delta_terms=None):
        """gmm_params: list of dicts(weights, means_invvars, inv_vars, gconsts) in scoring order.
        delta_terms: fp16 product terms of the (slot m - slot 0) part of the contraction when all slots share their
        variances -- 1, 2, 3, or None / 0 for the automatic choice (env FAKEBOB_GMM_DELTA_TERMS overrides None)."""
        self.lib = _lib.load()
        self.device = default_device() if device is None else device
        self.cfg = feat_cfg or FeatureConfig()
        self.cfg.check_supported()
        h = C.c_void_p()
        _lib.check(self.lib.fb_ctx_create(self.device, C.byref(h)))
        self.h = h
        fc = _lib.FeatConfig(self.cfg.sample_frequency, self.cfg.low_freq, self.cfg.high_freq, self.cfg.num_mel_bins,
                             self.cfg.num_ceps, self.cfg.preemph, self.cfg.cepstral_lifter,
                             self.cfg.vad_energy_threshold, self.cfg.vad_energy_mean_scale,
                             self.cfg.vad_proportion_threshold, self.cfg.vad_frames_context, self.cfg.cmn_window)
        _lib.check(self.lib.fb_set_feature_config(self.h, C.byref(fc)))
        self.n_models = len(gmm_params)
        self._params = gmm_params
        for slot, g in enumerate(gmm_params):
            w = np.ascontiguousarray(g["weights"], dtype=np.float32)
            miv = np.ascontiguousarray(g["means_invvars"], dtype=np.float32)
            iv = np.ascontiguousarray(g["inv_vars"], dtype=np.float32)
            gc = np.ascontiguousarray(g["gconsts"], dtype=np.float32)
            Cn, D = miv.shape
            _lib.check(self.lib.fb_load_diag_gmm(self.h, slot, _ptr(w), _ptr(miv), _ptr(iv), _ptr(gc), Cn, D))
        if delta_terms is None:
            delta_terms = int(os.environ.get("FAKEBOB_GMM_DELTA_TERMS", "0"))
        _lib.check(self.lib.fb_set_gmm_delta_terms(self.h, int(delta_terms)))
        _lib.check(self.lib.fb_finalize_gmms(self.h, self.n_models))
        self._kaldi_exact_from_env()
        self._nes_keep = None
        self._last_B = 0

    @classmethod
    def from_files(cls, paths, feat_cfg=None, device=None, delta_terms=None):
        return cls([kaldi_io.read_diag_gmm(p) for p in paths], feat_cfg=feat_cfg, device=device, delta_terms=delta_terms)

    def close(self):
        if getattr(self, "h", None):
            self.lib.fb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- scoring ---------------------------------------------------------------------------------
    def score_avg_ll(self, audio_list):
        """list of int16 1-D arrays -> (B, n_models) float64 average frame log-likelihoods."""
        B = len(audio_list)
        if B == 0:
            return np.zeros((0, self.n_models))
        lens = np.array([a.shape[0] for a in audio_list], dtype=np.int64)
        offsets = np.zeros(B + 1, dtype=np.int64)
        np.cumsum(lens, out=offsets[1:])
        wave = np.concatenate(audio_list) if B > 1 else audio_list[0]
        wave = np.ascontiguousarray(wave, dtype=np.int16)
        out = np.empty((B, self.n_models), dtype=np.float64)
        _lib.check(self.lib.fb_score_gmm_host(self.h, _ptr(wave), _ptr(offsets), B, _ptr(out)))
        self._last_B = B
        self._last_offsets = offsets
        return out

    def map_adapt(self, audio_list, mean_tau=10.0):
        """MAP mean-only adaptation of the (single) resident GMM -- the UBM -- to the pooled voiced frames of the
        audios: gmm-global-acc-stats | gmm-global-est-map --update-flags=m (build_spk_models.py:202-219).
        Returns dict(weights, means_invvars, inv_vars, gconsts, occupancy)."""
        if self.n_models != 1:
            raise ValueError("map_adapt needs an engine holding the UBM alone")
        B = len(audio_list)
        lens = np.array([a.shape[0] for a in audio_list], dtype=np.int64)
        offsets = np.zeros(B + 1, dtype=np.int64)
        np.cumsum(lens, out=offsets[1:])
        wave = np.ascontiguousarray(np.concatenate(audio_list) if B > 1 else audio_list[0], dtype=np.int16)
        g = self._params[0]
        Cn, D = g["means_invvars"].shape
        miv = np.empty((Cn, D), dtype=np.float32)
        gc = np.empty(Cn, dtype=np.float32)
        occ = np.empty(Cn, dtype=np.float64)
        _lib.check(self.lib.fb_map_adapt_host(self.h, _ptr(wave), _ptr(offsets), B, float(mean_tau), _ptr(miv), _ptr(gc), _ptr(occ)))
        self._last_B = B
        self._last_offsets = offsets
        return {"weights": np.array(g["weights"], dtype=np.float32), "means_invvars": miv,
                "inv_vars": np.array(g["inv_vars"], dtype=np.float32), "gconsts": gc, "occupancy": occ}

    def set_kaldi_exact(self, compress=False, text=False):
        """Non-ideal effects of the reference's real Kaldi path: CompressedMatrix round trip of the MFCCs
        (copy-feats --compress=true) and 7-significant-digit text round trips of scores / i-vectors.  Off by default;
        the scorers switch them on from FAKEBOB_KALDI_EXACT="compress,text"."""
        _lib.check(self.lib.fb_set_kaldi_exact(self.h, 1 if compress else 0, 1 if text else 0))

    def _kaldi_exact_from_env(self):
        opt = os.environ.get("FAKEBOB_KALDI_EXACT", "")
        if opt:
            o = {x.strip() for x in opt.replace("1", "compress,text").split(",")}
            self.set_kaldi_exact("compress" in o, "text" in o)

    def set_debug(self, on=True):
        _lib.check(self.lib.fb_set_debug(self.h, 1 if on else 0))

    def gmm_info(self):
        """-> dict(shared_variances, delta_terms, err_estimate) of the resident model image."""
        sh, dt, e = C.c_int(0), C.c_int(0), C.c_double(0)
        _lib.check(self.lib.fb_get_gmm_info(self.h, C.byref(sh), C.byref(dt), C.byref(e)))
        return {"shared_variances": bool(sh.value), "delta_terms": dt.value, "err_estimate": e.value}

    def set_delta_terms(self, terms):
        """Rebuild the resident model image with another number of difference terms (0 = automatic)."""
        _lib.check(self.lib.fb_set_gmm_delta_terms(self.h, int(terms)))
        _lib.check(self.lib.fb_finalize_gmms(self.h, self.n_models))

    def last_stages(self):
        """Per-stage outputs of the last score call (parity tests): dict(frames, voiced, mfcc, vad, feats, frame_ll)."""
        B = self._last_B
        frames = np.zeros(B, dtype=np.int32)
        voiced = np.zeros(B, dtype=np.int32)
        _lib.check(self.lib.fb_get_num_frames(self.h, B, _ptr(frames), _ptr(voiced)))
        T = int(frames.sum())
        mf = np.empty((T, 24), dtype=np.float32)
        _lib.check(self.lib.fb_get_mfcc(self.h, _ptr(mf), mf.size))
        vad = np.empty(T, dtype=np.int32)
        _lib.check(self.lib.fb_get_vad(self.h, _ptr(vad), vad.size))
        rows = int(voiced.sum())
        out = {"frames": frames, "voiced": voiced, "mfcc": mf, "vad": vad}
        fl = np.empty((self.n_models, rows), dtype=np.float32)
        _lib.check(self.lib.fb_get_frame_loglikes(self.h, _ptr(fl), fl.size))
        out["frame_ll"] = fl
        try:
            ft = np.empty((rows, 72), dtype=np.float32)
            _lib.check(self.lib.fb_get_features(self.h, _ptr(ft), ft.size))
            out["feats"] = ft
        except _lib.FakebobLibraryError:
            out["feats"] = None
        return out

    def features(self, wave_int16):
        """(Tv, 72) float32 front-end output for one utterance (used to build synthetic models on the GPU)."""
        self.set_debug(True)
        self.score_avg_ll([np.ascontiguousarray(wave_int16, dtype=np.int16)])
        return self.last_stages()["feats"]

    # ---- NES -------------------------------------------------------------------------------------
    def nes_init(self, audio, task, attack_type, n_speakers, label, threshold, adver_thresh, epsilon, max_iter,
                 max_lr, min_lr, samples_per_draw, sigma, momentum, plateau_length, plateau_drop,
                 rng="philox", seed=0, draw_base=0, z_means=None, z_stds=None, external=False):
        audio = np.ascontiguousarray(np.asarray(audio, dtype=np.float64).reshape(-1))
        p = _lib.NesParams()
        p.task = _lib.TASK[task]
        p.targeted = 1 if attack_type == "targeted" else 0
        p.label = -1 if label is None else int(label)
        p.n_speakers = n_speakers
        p.samples_per_draw = samples_per_draw
        p.max_iter = max_iter
        p.rng = _lib.RNG[rng]
        p.plateau_length = plateau_length
        p.threshold, p.adver_thresh, p.epsilon, p.sigma = threshold, adver_thresh, epsilon, sigma
        p.max_lr, p.min_lr, p.momentum, p.plateau_drop = max_lr, min_lr, momentum, plateau_drop
        p.seed, p.draw_base = seed, draw_base
        p.external_scorer = 1 if external else 0
        keep = [audio]
        if z_means is not None:
            zm = np.ascontiguousarray(z_means, dtype=np.float64)
            zs = np.ascontiguousarray(z_stds, dtype=np.float64)
            p.z_norm_means = zm.ctypes.data_as(C.POINTER(C.c_double))
            p.z_norm_stds = zs.ctypes.data_as(C.POINTER(C.c_double))
            keep += [zm, zs]
        _lib.check(self.lib.fb_nes_init(self.h, C.byref(p), _ptr(audio), audio.shape[0]))
        self._nes_keep = keep
        self._nes_N = audio.shape[0]
        self._nes_K = n_speakers
        self._nes_S2 = samples_per_draw // 2

    # ---- black-box scorers: the device keeps the attack state, the caller scores the batch ------------------------------
    def nes_ext_perturb(self):
        """-> (S+1, N) int16: row 0 the current adversarial audio, then +noise rows, then -noise rows (FAKEBOB.py:234-237)."""
        w = np.empty((2 * self._nes_S2 + 1, self._nes_N), dtype=np.int16)
        _lib.check(self.lib.fb_nes_ext_perturb(self.h, _ptr(w)))
        return w

    def nes_ext_update(self, scores, gradient_only=False):
        sc = np.ascontiguousarray(np.asarray(scores, dtype=np.float64).reshape(2 * self._nes_S2 + 1, self._nes_K))
        _lib.check(self.lib.fb_nes_ext_update(self.h, _ptr(sc), 1 if gradient_only else 0))

    def nes_gest(self):
        """-> (gradient estimate (N,), losses (S+1,), clean scores (K,)) of the last gradient-only step."""
        g = np.empty(self._nes_N, dtype=np.float64)
        ls = np.empty(2 * self._nes_S2 + 1, dtype=np.float64)
        sc = np.empty(self._nes_K, dtype=np.float64)
        _lib.check(self.lib.fb_nes_read_gest(self.h, _ptr(g), g.shape[0], _ptr(ls), _ptr(sc)))
        return g, ls, sc

    def nes_run(self, n_iters, noise=None):
        """noise (host rng): (n_iters, S/2, N) float64, pair-major."""
        if noise is not None:
            noise = np.ascontiguousarray(noise, dtype=np.float64)
            assert noise.shape == (n_iters, self._nes_S2, self._nes_N)
        _lib.check(self.lib.fb_nes_run(self.h, n_iters, _ptr(noise) if noise is not None else None))

    def nes_status(self):
        """-> (iterations done, stop code: 0 running, 1 early stop / candidate threshold reached, 2 accepted (estimate mode))."""
        it, st = C.c_int(0), C.c_int(0)
        _lib.check(self.lib.fb_nes_status(self.h, C.byref(it), C.byref(st)))
        return it.value, st.value

    def nes_estimate_begin(self, accept_threshold):
        _lib.check(self.lib.fb_nes_estimate_begin(self.h, float(accept_threshold)))

    def nes_continue(self, threshold):
        _lib.check(self.lib.fb_nes_continue(self.h, float(threshold)))

    def nes_log(self, max_rows):
        rows = np.zeros((max_rows, 4 + self._nes_K), dtype=np.float64)
        n = _lib.check(self.lib.fb_nes_read_log(self.h, _ptr(rows), max_rows))
        return rows[:n]

    def nes_adver(self):
        a = np.empty(self._nes_N, dtype=np.float64)
        _lib.check(self.lib.fb_nes_read_adver(self.h, _ptr(a), a.shape[0]))
        return a

    def nes_grad(self):
        a = np.empty(self._nes_N, dtype=np.float64)
        _lib.check(self.lib.fb_nes_read_grad(self.h, _ptr(a), a.shape[0]))
        return a

    def nes_set_threshold(self, threshold):
        _lib.check(self.lib.fb_nes_set_threshold(self.h, float(threshold)))

    def nes_get_grad(self, noise=None):
        """-> (final_loss, grad (N,), adver_loss, score0 (K,)).  noise (host rng): (S/2, N) float64."""
        if noise is not None:
            noise = np.ascontiguousarray(noise, dtype=np.float64)
            assert noise.shape == (self._nes_S2, self._nes_N)
        fl, al = C.c_double(0), C.c_double(0)
        sc = np.zeros(self._nes_K, dtype=np.float64)
        g = np.empty(self._nes_N, dtype=np.float64)
        _lib.check(self.lib.fb_nes_get_grad(self.h, _ptr(noise) if noise is not None else None,
                                            C.byref(fl), C.byref(al), _ptr(sc), _ptr(g)))
        return fl.value, g, al.value, sc

    def nes_apply_update(self, lr):
        _lib.check(self.lib.fb_nes_apply_update(self.h, float(lr)))

    def kernel_launches(self):
        n = C.c_int64(0)
        _lib.check(self.lib.fb_nes_kernel_launches(self.h, C.byref(n)))
        return n.value

    STAGES = ("perturb", "mfcc", "vad_scan", "feats", "gmm", "gmm_reduce", "loss", "update",
              "gselect", "fgmm_post", "ivec_stats", "ivec_lin", "ivec_quad", "ivec_solve", "plda", "spare")

    def profile(self, on=True):
        _lib.check(self.lib.fb_profile_enable(self.h, 1 if on else 0))

    def profile_read(self):
        """-> {stage: (total_ms, launches)} accumulated since profile(True)."""
        ms = np.zeros(16, dtype=np.float64)
        cnt = np.zeros(16, dtype=np.int64)
        _lib.check(self.lib.fb_profile_read(self.h, _ptr(ms), _ptr(cnt)))
        return {s: (float(ms[i]), int(cnt[i])) for i, s in enumerate(self.STAGES)}

    def voiced_rows(self):
        r = C.c_int(0)
        _lib.check(self.lib.fb_get_voiced_rows(self.h, C.byref(r)))
        return r.value

    def set_stream(self, cuda_stream_ptr):
        _lib.check(self.lib.fb_set_stream(self.h, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    def synchronize(self):
        _lib.check(self.lib.fb_synchronize(self.h))

    # ---- multi-GPU ---------------------------------------------------------------------------------
    def comm_init_from_torch(self, max_samples=1 << 20):
        """Create the NCCL communicator for this engine using torch.distributed for the id exchange."""
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        if world == 1:
            return
        buf = (C.c_char * 128)()
        if rank == 0:
            _lib.check(self.lib.fb_comm_unique_id(buf))
        obj = [bytes(buf)]
        dist.broadcast_object_list(obj, src=0)
        ident = (C.c_char * 128).from_buffer_copy(obj[0])
        _lib.check(self.lib.fb_comm_init(self.h, ident, rank, world))
        # exchange buffers in peer memory (one node, <= 8 ranks): the gradient partials are then pulled over NVLink inside
        # the update kernel; if the mapping is refused the ncclAllReduce path stays in use
        self.p2p = False
        if world <= 8 and os.environ.get("FB_NO_P2P") is None:
            hbuf = (C.c_char * 64)()
            _lib.check(self.lib.fb_comm_p2p_export(self.h, int(max_samples), hbuf))
            handles = [None] * world
            dist.all_gather_object(handles, bytes(hbuf))
            allh = (C.c_char * (64 * world)).from_buffer_copy(b"".join(handles))
            ok = self.lib.fb_comm_p2p_import(self.h, allh) == 0
            flags = [None] * world
            dist.all_gather_object(flags, ok)
            self.p2p = all(flags)
            if not self.p2p:
                os.environ["FB_NO_P2P"] = "1"          # every rank must take the same path


class ExternalEngine(GmmEngine):
    """A bare context for attacking a black-box scorer (reference README.md:136): no resident models; the NES state, noise
    generator, quantisation, gradient and update run on the device, the caller scores each batch."""

    def __init__(self, device=None):
        self.lib = _lib.load()
        self.device = default_device() if device is None else device
        h = C.c_void_p()
        _lib.check(self.lib.fb_ctx_create(self.device, C.byref(h)))
        self.h = h
        self.n_models = 0
        self._nes_keep = None
        self._last_B = 0


class IvectorEngine(GmmEngine):
    """Resident full UBM + i-vector extractor + LDA/PLDA back-end on one B200 (replaces ivector_PLDA_kaldiHelper)."""

    def __init__(self, pre_model_dir, feat_cfg=None, device=None):
        self.lib = _lib.load()
        self.device = default_device() if device is None else device
        self.cfg = feat_cfg or FeatureConfig()
        self.cfg.check_supported()
        h = C.c_void_p()
        _lib.check(self.lib.fb_ctx_create(self.device, C.byref(h)))
        self.h = h
        fc = _lib.FeatConfig(self.cfg.sample_frequency, self.cfg.low_freq, self.cfg.high_freq, self.cfg.num_mel_bins,
                             self.cfg.num_ceps, self.cfg.preemph, self.cfg.cepstral_lifter,
                             self.cfg.vad_energy_threshold, self.cfg.vad_energy_mean_scale,
                             self.cfg.vad_proportion_threshold, self.cfg.vad_frames_context, self.cfg.cmn_window)
        _lib.check(self.lib.fb_set_feature_config(self.h, C.byref(fc)))
        pre = pre_model_dir
        fg = kaldi_io.read_full_gmm(os.path.join(pre, "final.ubm"))
        w = np.ascontiguousarray(fg["weights"], dtype=np.float32)
        mic = np.ascontiguousarray(fg["means_invcovars"], dtype=np.float32)
        ic = np.ascontiguousarray(fg["inv_covars"], dtype=np.float32)
        gc = np.ascontiguousarray(fg["gconsts"], dtype=np.float32)
        Cn, D = mic.shape
        self._C = Cn
        _lib.check(self.lib.fb_load_full_gmm(self.h, _ptr(w), _ptr(mic), _ptr(ic), _ptr(gc), Cn, D))
        ie = kaldi_io.read_ivector_extractor(os.path.join(pre, "final.ie"))
        if ie["w"].size:
            raise ValueError("i-vector extractors with weight projection (<w> non-empty) are not supported")
        M = np.ascontiguousarray(ie["M"], dtype=np.float64)
        S = np.ascontiguousarray(ie["sigma_inv"], dtype=np.float64)
        self.R = M.shape[2]
        _lib.check(self.lib.fb_load_ivector_extractor(self.h, _ptr(M), _ptr(S), float(ie["prior_offset"]), Cn, D, self.R))
        mean_vec = np.ascontiguousarray(kaldi_io.read_vector(os.path.join(pre, "mean.vec")), dtype=np.float32)
        tm = np.ascontiguousarray(kaldi_io.read_matrix(os.path.join(pre, "transform.mat")), dtype=np.float32)
        pl = kaldi_io.read_plda(os.path.join(pre, "plda"))
        self.L = tm.shape[0]
        _lib.check(self.lib.fb_load_plda_backend(self.h, _ptr(mean_vec), _ptr(tm), tm.shape[1],
                                                 _ptr(np.ascontiguousarray(pl["mean"])), _ptr(np.ascontiguousarray(pl["transform"])),
                                                 _ptr(np.ascontiguousarray(pl["psi"])), self.R, self.L))
        self.n_models = 1
        self.K = 0
        self._kaldi_exact_from_env()
        self._nes_keep = None
        self._last_B = 0

    def set_enrolled(self, ivectors):
        e = np.ascontiguousarray(np.atleast_2d(ivectors), dtype=np.float32)
        _lib.check(self.lib.fb_set_enrolled_ivectors(self.h, _ptr(e), e.shape[0]))
        self.K = e.shape[0]

    def _pack(self, audio_list):
        B = len(audio_list)
        lens = np.array([a.shape[0] for a in audio_list], dtype=np.int64)
        offsets = np.zeros(B + 1, dtype=np.int64)
        np.cumsum(lens, out=offsets[1:])
        wave = np.ascontiguousarray(np.concatenate(audio_list) if B > 1 else audio_list[0], dtype=np.int16)
        return wave, offsets, B

    def extract_ivectors(self, audio_list):
        """list of int16 arrays -> (B, R) float32 raw i-vectors (what ivector-extract writes)."""
        wave, offsets, B = self._pack(audio_list)
        out = np.empty((B, self.R), dtype=np.float32)
        _lib.check(self.lib.fb_score_ivector_host(self.h, _ptr(wave), _ptr(offsets), B, None, _ptr(out)))
        self._last_B = B
        return out

    def score_plda(self, audio_list, want_ivectors=False):
        """-> (B, K) float64 PLDA log-likelihood ratios against the enrolled speakers."""
        wave, offsets, B = self._pack(audio_list)
        out = np.empty((B, self.K), dtype=np.float64)
        iv = np.empty((B, self.R), dtype=np.float32) if want_ivectors else None
        _lib.check(self.lib.fb_score_ivector_host(self.h, _ptr(wave), _ptr(offsets), B, _ptr(out), _ptr(iv) if iv is not None else None))
        self._last_B = B
        return (out, iv) if want_ivectors else out

    def stats(self, b):
        """Intermediate results of utterance b of the last batch (parity tests):
        dict(gamma (C,), X (C,72), lin (R,), quad (R(R+1)/2,) packed lower triangle), float64."""
        Cn = self._C
        out = {"gamma": np.empty(Cn), "X": np.empty((Cn, 72)), "lin": np.empty(self.R), "quad": np.empty(self.R * (self.R + 1) // 2)}
        _lib.check(self.lib.fb_get_ivector_stats(self.h, int(b), _ptr(out["gamma"]), _ptr(out["X"]), _ptr(out["lin"]), _ptr(out["quad"])))
        return out

    def posteriors(self):
        """Gaussian selection and pruned posteriors of the last batch: (rows, 20) int32 / float32."""
        rows = self.voiced_rows()
        g = np.empty((rows, 20), dtype=np.int32)
        p = np.empty((rows, 20), dtype=np.float32)
        _lib.check(self.lib.fb_get_posteriors(self.h, _ptr(g), _ptr(p), rows))
        return g, p
