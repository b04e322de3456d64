"""fakebob_b200: B200-native (sm_100a) implementation of FAKEBOB's NES attack hot path.

Public API mirrors the reference: FAKEBOB.FakeBob, gmm_ubm_{CSI,OSI,SV}.gmm_*, ivector_PLDA_{CSI,OSI,SV}.iv_*.
Put ``fakebob_b200/dropin`` first on ``sys.path`` and the reference's attackMain.py
runs unmodified on the GPU path.
"""
import os

__version__ = "0.1.0"
DROPIN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")
