"""Drop-in module name of the reference's gmm_ubm_CSI.py (attackMain.py:15-21 imports gmm_CSI from it)."""
from .gmm_scorers import gmm_CSI  # noqa: F401
