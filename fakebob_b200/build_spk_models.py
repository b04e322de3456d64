"""Speaker enrolment on the B200: the device counterpart of the reference's ``build_spk_models.py``.

The reference script (``/root/reference/build_spk_models.py``) enrols every wav of ``data/enrollment-set`` twice:
  step 1 (:103-152)  i-vector identity: ``sid/extract_ivectors.sh`` on enrolment + z-norm audio, PLDA scores of the z-norm
                     cohort against every enrolled i-vector -> z-norm mean / std, pickle ``model/<spk>.iv``;
  step 2 (:161-277)  GMM identity: per speaker ``gmm-global-acc-stats --update-flags=m`` + ``gmm-global-est-map
                     --update-flags=m`` (MAP mean-only adaptation of final.dubm, tau = 10) -> ``model/<spk>-identity.gmm``,
                     average log-likelihoods of the z-norm cohort under every identity GMM -> z-norm mean / std,
                     pickle ``model/<spk>.gmm``.
Here both steps run through the C-ABI (``fb_map_adapt_host``, ``fb_score_gmm_host``, ``fb_score_ivector_host``): no Kaldi
processes, no feature archives.  Outputs have the reference's formats: Kaldi binary DiagGmm files, a Kaldi text ark of
i-vectors addressed by ``<ark>:<offset>`` scp targets, and pickled 5-lists
``[spk_id, utt_id, identity_location (absolute), z_norm_mean, z_norm_std]`` (build_spk_models.py:15-30).

Utterance / speaker ids follow the reference: ``utt_id`` = file name up to the first ".", ``spk_id`` = ``utt_id`` up to
the first "-" (build_spk_models.py:81-83).
"""
import os
import pickle

import numpy as np

from . import kaldi_io
from .config import load_feature_config
from .engine import GmmEngine, IvectorEngine, to_audio_list


def list_audio_dir(path):
    """-> (utt_ids, spk_ids, paths) like build_spk_models.py:76-99 (directory order made deterministic by sorting)."""
    utt, spk, paths = [], [], []
    for name in sorted(os.listdir(path)):
        u = name.split(".")[0]
        utt.append(u)
        spk.append(u.split("-")[0])
        paths.append(os.path.join(path, name))
    return utt, spk, paths


def read_wav_int16(path):
    from scipy.io import wavfile
    fs, a = wavfile.read(path)
    if fs != 16000:
        raise ValueError("%s: sampling rate %d, the models need 16000" % (path, fs))
    a = np.asarray(a)
    if a.ndim > 1:
        a = a[:, 0]
    if a.dtype != np.int16:
        raise ValueError("%s: expected 16-bit PCM" % path)
    return np.ascontiguousarray(a)


def enroll_gmm(enroll_audios, enroll_spk_ids, enroll_utt_ids, z_norm_audios, pre_model_dir, model_dir, mean_tau=10.0,
               device=None):
    """Step 2 of the reference script.  Audios are int16 arrays (or anything ``to_audio_list`` accepts).
    Returns the list of speaker models (5-lists) in enrolment order."""
    pre_model_dir = os.path.abspath(pre_model_dir)
    model_dir = os.path.abspath(model_dir)
    os.makedirs(model_dir, exist_ok=True)
    cfg = load_feature_config(pre_model_dir)
    ubm = kaldi_io.read_diag_gmm(os.path.join(pre_model_dir, "final.dubm"))
    eng = GmmEngine([ubm], feat_cfg=cfg, device=device)
    identities, paths = [], []
    try:
        for audio, spk_id in zip(to_audio_list(list(enroll_audios)), enroll_spk_ids):
            g = eng.map_adapt([audio], mean_tau=mean_tau)                      # one utterance per speaker (:188-193)
            path = os.path.join(model_dir, spk_id + "-identity.gmm")
            kaldi_io.write_diag_gmm(path, g["weights"], g["means_invvars"], g["inv_vars"], g["gconsts"])
            identities.append(g)
            paths.append(path)
    finally:
        eng.close()
    # z-norm: average log-likelihood of every cohort utterance under every identity GMM (:259-261), 32 models per pass
    z_list = to_audio_list(list(z_norm_audios))
    cols = []
    for i in range(0, len(identities), 32):
        sc = GmmEngine(identities[i:i + 32], feat_cfg=cfg, device=device)
        try:
            cols.append(sc.score_avg_ll(z_list))
        finally:
            sc.close()
    score_array = np.concatenate(cols, axis=1)
    z_mean = np.mean(score_array, axis=0).flatten()
    z_std = np.std(score_array, axis=0).flatten()
    models = []
    for i, spk_id in enumerate(enroll_spk_ids):
        m = [spk_id, enroll_utt_ids[i], os.path.abspath(paths[i]), float(z_mean[i]), float(z_std[i])]
        with open(os.path.join(model_dir, spk_id + ".gmm"), "wb") as f:
            pickle.dump(m, f, protocol=-1)
        models.append(m)
    return models


def enroll_ivector(enroll_audios, enroll_spk_ids, enroll_utt_ids, z_norm_audios, pre_model_dir, model_dir, ivector_dir=None,
                   device=None):
    """Step 1 of the reference script: raw i-vector identities (text ark + scp targets, 'ark,t' like
    ivector_PLDA_kaldiHelper.py:202-211) and PLDA z-norm statistics of the cohort against every enrolled speaker."""
    pre_model_dir = os.path.abspath(pre_model_dir)
    model_dir = os.path.abspath(model_dir)
    ivector_dir = os.path.abspath(ivector_dir or os.path.join(model_dir, "ivector-build-model-iv"))
    os.makedirs(model_dir, exist_ok=True)
    os.makedirs(ivector_dir, exist_ok=True)
    eng = IvectorEngine(pre_model_dir, feat_cfg=load_feature_config(pre_model_dir), device=device)
    try:
        ivs = eng.extract_ivectors(to_audio_list(list(enroll_audios)))
        targets = kaldi_io.write_text_vector_ark(os.path.join(ivector_dir, "ivector.1.ark"),
                                                 list(zip(enroll_utt_ids, [np.asarray(v, dtype=np.float64) for v in ivs])))
        with open(os.path.join(ivector_dir, "ivector.scp"), "w") as f:
            for u in enroll_utt_ids:
                f.write("%s %s\n" % (u, targets[u]))
        # the scorers read the identities back from the text ark (7 significant digits): score against those
        enrolled = np.stack([np.asarray(kaldi_io.read_vector(targets[u]), dtype=np.float32) for u in enroll_utt_ids])
        eng.set_enrolled(enrolled)
        scores = eng.score_plda(to_audio_list(list(z_norm_audios)))          # (n_cohort, K)
    finally:
        eng.close()
    models = []
    for i, spk_id in enumerate(enroll_spk_ids):
        m = [spk_id, enroll_utt_ids[i], os.path.abspath(targets[enroll_utt_ids[i]]), float(np.mean(scores[:, i])),
             float(np.std(scores[:, i]))]
        with open(os.path.join(model_dir, spk_id + ".iv"), "wb") as f:
            pickle.dump(m, f, protocol=-1)
        models.append(m)
    return models


def main(enroll_dir="./data/enrollment-set", z_norm_dir="./data/z-norm-set", pre_model_dir="./pre-models", model_dir="./model",
         archs=("iv", "gmm"), device=None):
    """Same defaults and outputs as running the reference's build_spk_models.py from its working directory."""
    e_utt, e_spk, e_paths = list_audio_dir(enroll_dir)
    _, _, z_paths = list_audio_dir(z_norm_dir)
    e_audio = [read_wav_int16(p) for p in e_paths]
    z_audio = [read_wav_int16(p) for p in z_paths]
    out = {}
    if "iv" in archs:
        print("----- step 1: generate ivector identity and corresponding speaker model -----")
        out["iv"] = enroll_ivector(e_audio, e_spk, e_utt, z_audio, pre_model_dir, model_dir, device=device)
    if "gmm" in archs:
        print("----- step 2: generate gmm identity and corresponding speaker model -----")
        out["gmm"] = enroll_gmm(e_audio, e_spk, e_utt, z_audio, pre_model_dir, model_dir, device=device)
    for models in out.values():
        for m in models:
            print(m)
    return out


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--enroll-dir", default="./data/enrollment-set")
    ap.add_argument("--z-norm-dir", default="./data/z-norm-set")
    ap.add_argument("--pre-model-dir", default="./pre-models")
    ap.add_argument("--model-dir", default="./model")
    ap.add_argument("--archs", default="iv,gmm")
    a = ap.parse_args()
    main(a.enroll_dir, a.z_norm_dir, a.pre_model_dir, a.model_dir, tuple(a.archs.split(",")))
