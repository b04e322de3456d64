"""Drop-in module name of the reference's ivector_PLDA_OSI.py (attackMain.py:15-21 imports iv_OSI from it)."""
from .iv_scorers import iv_OSI  # noqa: F401
