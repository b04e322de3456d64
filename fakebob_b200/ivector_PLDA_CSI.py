"""Drop-in module name of the reference's ivector_PLDA_CSI.py (attackMain.py:15-21 imports iv_CSI from it)."""
from .iv_scorers import iv_CSI  # noqa: F401
