"""Top-level alias so `from gmm_ubm_OSI import gmm_OSI` (attackMain.py:15-21) resolves to the B200 build."""
from fakebob_b200.gmm_scorers import gmm_OSI  # noqa: F401
