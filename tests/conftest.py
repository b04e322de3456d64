import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no GPU in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_oracle_gmm(path):
    from oracle import kaldi_files      # the oracle's own file reader, not the product's
    from oracle.diag_gmm import DiagGmm
    g = kaldi_files.read_diag_gmm(path)
    return DiagGmm(g["weights"], g["means_invvars"], g["inv_vars"], g["gconsts"])


@pytest.fixture(scope="session")
def small_tree(tmp_path_factory):
    """Synthetic pre-models/ + model/ tree: 256-mixture UBM, 3 speakers, 2 s utterances (oracle features)."""
    from fakebob_b200 import synth
    from oracle import kaldi_feats as kf
    root = str(tmp_path_factory.mktemp("small_tree"))
    tree = synth.build_gmm_tree(root, kf.voiced_features, n_speakers=3, C=256, n_ubm_utts=12, n_samples=32000,
                                n_znorm_utts=4, em_iters=2)
    tree["root"] = root
    return tree


@pytest.fixture(scope="session")
def small_oracle_models(small_tree):
    ubm = load_oracle_gmm(small_tree["ubm"])
    spk = [load_oracle_gmm(m[2]) for m in small_tree["models"]]
    return ubm, spk


def test_audio(seed, spk=0, n=32000):
    from fakebob_b200 import synth
    return synth.synth_utterance(seed=seed, spk_seed=spk, n_samples=n)


@pytest.fixture(scope="session")
def small_iv_tree(small_tree):
    """Adds final.ubm / final.ie / mean.vec / transform.mat / plda and enrolled speakers (*.iv pickles) to small_tree."""
    from fakebob_b200 import synth
    from oracle.ivector import load_system
    root = small_tree["root"]
    synth.build_ivector_params(root, small_tree["ubm_params"], R=40, L=20)
    system = load_system(small_tree["pre_model_dir"])
    spk = synth.build_ivector_speakers(root, system.extract, system.plda_scores, n_speakers=3, n_samples=32000, n_znorm_utts=4)
    out = dict(small_tree)
    out.update(iv_models=spk["models"], enrolled=spk["enrolled"], system=system)
    return out
