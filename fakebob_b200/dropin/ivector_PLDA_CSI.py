"""Top-level alias so `from ivector_PLDA_CSI import iv_CSI` (attackMain.py:15-21) resolves to the B200 build."""
from fakebob_b200.iv_scorers import iv_CSI  # noqa: F401
