"""Top-level alias so `from gmm_ubm_CSI import gmm_CSI` (attackMain.py:15-21) resolves to the B200 build."""
from fakebob_b200.gmm_scorers import gmm_CSI  # noqa: F401
