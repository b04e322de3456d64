"""Evaluation helpers: what the reference's ``test.py`` computes around the scorers.

``test.py`` itself runs unmodified against the drop-in modules (it imports the scorers by module name, test.py:8-13);
these functions give the same numbers to callers that hold audio in memory.
"""
import numpy as np


def set_threshold(score_target, score_untarget):
    """Equal-error-rate threshold search of test.py:46-71: among the TARGET scores taken as candidate thresholds, in
    their given order, the first one minimising |FRR - FAR| with FRR = %(target < thr), FAR = %(untarget >= thr).
    Returns (threshold, frr, far) in percent like the reference.  O(n log n) instead of the reference's O(n^2)."""
    st = np.asarray(score_target, dtype=np.float64).reshape(-1)
    su = np.asarray(score_untarget, dtype=np.float64).reshape(-1)
    if st.size == 0 or su.size == 0:
        raise ValueError("set_threshold needs at least one target and one untarget score")
    st_sorted = np.sort(st)
    su_sorted = np.sort(su)
    frr = np.searchsorted(st_sorted, st, side="left") * 100 / st.size              # targets strictly below the candidate
    far = (su.size - np.searchsorted(su_sorted, st, side="left")) * 100 / su.size  # untargets at or above it
    diff = np.abs(frr - far)
    i = int(np.argmin(diff))                                                       # first minimum, like the strict '<'
    return float(st[i]), float(frr[i]), float(far[i])


def csi_accuracy(model, audio_list, labels, **kw):
    """Closed-set identification accuracy in percent (test.py:76-99): decisions from ``make_decisions`` vs speaker
    indices into ``model.spk_ids``."""
    decisions, _ = model.make_decisions(audio_list, **kw)
    decisions = np.atleast_1d(np.asarray(decisions))
    labels = np.atleast_1d(np.asarray(labels))
    return float(np.count_nonzero(decisions == labels) * 100 / decisions.size)


def sv_error_rates(model, target_audio, illegal_audio, threshold=None, **kw):
    """Speaker verification (test.py SV sections): scores of the enrolled speaker's own audio and of impostors;
    with ``threshold=None`` the EER threshold is searched with :func:`set_threshold` and stored in ``model.threshold``.
    Returns dict(threshold, frr, far)."""
    st = np.atleast_1d(np.asarray(model.score(target_audio, **kw), dtype=np.float64))
    su = np.atleast_1d(np.asarray(model.score(illegal_audio, **kw), dtype=np.float64))
    if threshold is None:
        threshold, frr, far = set_threshold(st, su)
        model.threshold = threshold
    else:
        frr = float(np.count_nonzero(st < threshold) * 100 / st.size)
        far = float(np.count_nonzero(su >= threshold) * 100 / su.size)
    return {"threshold": float(threshold), "frr": frr, "far": far}


def osi_error_rates(model, audio_list, labels, illegal_audio, threshold=None, **kw):
    """Open-set identification (test.py OSI sections): FRR = enrolled audio rejected, IER = enrolled audio accepted as
    the wrong speaker, FAR = impostor audio accepted; the threshold search uses the maximum score per audio."""
    sc_t = np.atleast_2d(np.asarray(model.score(audio_list, **kw), dtype=np.float64))
    sc_u = np.atleast_2d(np.asarray(model.score(illegal_audio, **kw), dtype=np.float64))
    labels = np.atleast_1d(np.asarray(labels))
    if threshold is None:
        threshold, _, _ = set_threshold(sc_t.max(axis=1), sc_u.max(axis=1))
        model.threshold = threshold
    dec_t = np.where(sc_t.max(axis=1) < threshold, -1, sc_t.argmax(axis=1))
    dec_u = np.where(sc_u.max(axis=1) < threshold, -1, sc_u.argmax(axis=1))
    n = dec_t.size
    return {"threshold": float(threshold), "frr": float(np.count_nonzero(dec_t == -1) * 100 / n),
            "ier": float(np.count_nonzero((dec_t != -1) & (dec_t != labels)) * 100 / n),
            "far": float(np.count_nonzero(dec_u != -1) * 100 / dec_u.size)}
