"""Top-level alias so `from FAKEBOB import FakeBob` (attackMain.py:21) resolves to the B200 build."""
from fakebob_b200.FAKEBOB import FakeBob, UNTARGETED  # noqa: F401
