"""Drop-in module name of the reference's ivector_PLDA_SV.py (attackMain.py:15-21 imports iv_SV from it)."""
from .iv_scorers import iv_SV  # noqa: F401
