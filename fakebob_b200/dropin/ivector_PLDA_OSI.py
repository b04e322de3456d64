"""Top-level alias so `from ivector_PLDA_OSI import iv_OSI` (attackMain.py:15-21) resolves to the B200 build."""
from fakebob_b200.iv_scorers import iv_OSI  # noqa: F401
