"""Drop-in module name of the reference's gmm_ubm_OSI.py (attackMain.py:15-21 imports gmm_OSI from it)."""
from .gmm_scorers import gmm_OSI  # noqa: F401
