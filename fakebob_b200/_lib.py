"""ctypes binding of libfakebob_b200.so (include/fakebob_b200.h).

There is no CPU fallback: if the library is missing, or no sm_100 GPU is visible when a
context is created, the caller gets an exception.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FB_LIB_PATH") or os.path.join(_HERE, "libfakebob_b200.so")   # override: kernel experiments only


class FakebobLibraryError(RuntimeError):
    pass


class FeatConfig(C.Structure):
    _fields_ = [("sample_frequency", C.c_float), ("low_freq", C.c_float), ("high_freq", C.c_float),
                ("num_mel_bins", C.c_int), ("num_ceps", C.c_int), ("preemph", C.c_float),
                ("cepstral_lifter", C.c_float), ("vad_energy_threshold", C.c_float),
                ("vad_energy_mean_scale", C.c_float), ("vad_proportion_threshold", C.c_float),
                ("vad_frames_context", C.c_int), ("cmn_window", C.c_int)]


class NesParams(C.Structure):
    _fields_ = [("task", C.c_int), ("targeted", C.c_int), ("label", C.c_int), ("n_speakers", C.c_int),
                ("samples_per_draw", C.c_int), ("max_iter", C.c_int), ("rng", C.c_int), ("plateau_length", C.c_int),
                ("threshold", C.c_double), ("adver_thresh", C.c_double), ("epsilon", C.c_double),
                ("sigma", C.c_double), ("max_lr", C.c_double), ("min_lr", C.c_double), ("momentum", C.c_double),
                ("plateau_drop", C.c_double), ("seed", C.c_uint64), ("draw_base", C.c_uint64),
                ("z_norm_means", C.POINTER(C.c_double)), ("z_norm_stds", C.POINTER(C.c_double)), ("external_scorer", C.c_int)]


TASK = {"CSI": 0, "OSI": 1, "SV": 2}
RNG = {"numpy": 0, "host": 0, "philox": 1}

_P = C.c_void_p
_SIGS = {
    "fb_last_error": (C.c_char_p, []),
    "fb_version": (C.c_int, []),
    "fb_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "fb_ctx_destroy": (C.c_int, [_P]),
    "fb_set_stream": (C.c_int, [_P, _P]),
    "fb_synchronize": (C.c_int, [_P]),
    "fb_set_feature_config": (C.c_int, [_P, C.POINTER(FeatConfig)]),
    "fb_set_kaldi_exact": (C.c_int, [_P, C.c_int, C.c_int]),
    "fb_load_diag_gmm": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, C.c_int, C.c_int]),
    "fb_finalize_gmms": (C.c_int, [_P, C.c_int]),
    "fb_set_gmm_delta_terms": (C.c_int, [_P, C.c_int]),
    "fb_get_gmm_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "fb_score_gmm_host": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "fb_score_gmm_dev": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "fb_map_adapt_host": (C.c_int, [_P, _P, _P, C.c_int, C.c_double, _P, _P, _P]),
    "fb_load_full_gmm": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int]),
    "fb_load_ivector_extractor": (C.c_int, [_P, _P, _P, C.c_double, C.c_int, C.c_int, C.c_int]),
    "fb_load_plda_backend": (C.c_int, [_P, _P, _P, C.c_int, _P, _P, _P, C.c_int, C.c_int]),
    "fb_set_enrolled_ivectors": (C.c_int, [_P, _P, C.c_int]),
    "fb_score_ivector_host": (C.c_int, [_P, _P, _P, C.c_int, _P, _P]),
    "fb_get_posteriors": (C.c_int, [_P, _P, _P, C.c_int64]),
    "fb_get_ivector_stats": (C.c_int, [_P, C.c_int, _P, _P, _P, _P]),
    "fb_set_debug": (C.c_int, [_P, C.c_int]),
    "fb_get_num_frames": (C.c_int, [_P, C.c_int, _P, _P]),
    "fb_get_mfcc": (C.c_int, [_P, _P, C.c_int64]),
    "fb_get_vad": (C.c_int, [_P, _P, C.c_int64]),
    "fb_get_features": (C.c_int, [_P, _P, C.c_int64]),
    "fb_get_frame_loglikes": (C.c_int, [_P, _P, C.c_int64]),
    "fb_nes_init": (C.c_int, [_P, C.POINTER(NesParams), _P, C.c_int64]),
    "fb_nes_set_threshold": (C.c_int, [_P, C.c_double]),
    "fb_nes_run": (C.c_int, [_P, C.c_int, _P]),
    "fb_nes_status": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "fb_nes_read_log": (C.c_int, [_P, _P, C.c_int]),
    "fb_nes_read_adver": (C.c_int, [_P, _P, C.c_int64]),
    "fb_nes_read_grad": (C.c_int, [_P, _P, C.c_int64]),
    "fb_nes_estimate_begin": (C.c_int, [_P, C.c_double]),
    "fb_nes_continue": (C.c_int, [_P, C.c_double]),
    "fb_nes_ext_perturb": (C.c_int, [_P, _P]),
    "fb_nes_ext_update": (C.c_int, [_P, _P, C.c_int]),
    "fb_nes_read_gest": (C.c_int, [_P, _P, C.c_int64, _P, _P]),
    "fb_nes_get_grad": (C.c_int, [_P, _P, C.POINTER(C.c_double), C.POINTER(C.c_double), _P, _P]),
    "fb_nes_apply_update": (C.c_int, [_P, C.c_double]),
    "fb_nes_kernel_launches": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "fb_profile_enable": (C.c_int, [_P, C.c_int]),
    "fb_profile_read": (C.c_int, [_P, _P, _P]),
    "fb_get_voiced_rows": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "fb_comm_unique_id": (C.c_int, [_P]),
    "fb_comm_init": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "fb_comm_destroy": (C.c_int, [_P]),
    "fb_comm_p2p_export": (C.c_int, [_P, C.c_int64, _P]),
    "fb_comm_p2p_import": (C.c_int, [_P, _P]),
}

EXPORTS = tuple(_SIGS)
_lib = None


def load():
    """dlopen the library and attach prototypes.  Raises FakebobLibraryError if it was never built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FakebobLibraryError(
            "%s not found: build it with `python -m fakebob_b200.build` (needs nvcc, sm_100a). "
            "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    """Raise on a negative return code; non-negative values pass through (some calls return counts)."""
    if rc < 0:
        msg = load().fb_last_error().decode("utf-8", "replace")
        raise FakebobLibraryError("libfakebob_b200 error %d: %s" % (rc, msg))
    return rc
