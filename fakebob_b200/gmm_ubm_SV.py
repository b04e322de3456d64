"""Drop-in module name of the reference's gmm_ubm_SV.py (attackMain.py:15-21 imports gmm_SV from it)."""
from .gmm_scorers import gmm_SV  # noqa: F401
