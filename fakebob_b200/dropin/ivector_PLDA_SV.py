"""Top-level alias so `from ivector_PLDA_SV import iv_SV` (attackMain.py:15-21) resolves to the B200 build."""
from fakebob_b200.iv_scorers import iv_SV  # noqa: F401
