"""Top-level alias so `from gmm_ubm_SV import gmm_SV` (attackMain.py:15-21) resolves to the B200 build."""
from fakebob_b200.gmm_scorers import gmm_SV  # noqa: F401
