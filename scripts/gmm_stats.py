"""Reads the wait-time counters of the GMM_STATS kernel variant (scripts/build_variant.sh stats fb_gmm.cu -DGMM_STATS)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import runpy
runpy.run_path(os.path.join(os.path.dirname(__file__), "gmm_time.py"))
from fakebob_b200 import _lib
lib = _lib.load()
buf = np.zeros((148, 16), np.int64)
lib.fb_debug_gmm_stats.argtypes = [ctypes.c_void_p]
assert lib.fb_debug_gmm_stats(buf.ctypes.data) == 0
t0 = buf[:, 0].min()
print("kernel span (globaltimer, first entry -> last exit): %.1f us; entry skew max %.1f us" % (((buf[:, 0] + buf[:, 1]).max() - t0) / 1e3, (buf[:, 0].max() - t0) / 1e3))
names = ["exit-entry ns", "setup clk", "total clk", "producer total", "producer wait", "issuer0 total", "issuer0 wait", "issuer1 total", "issuer1 wait",
         "epi w3 total", "epi w3 wait", "epi w7 total", "epi w7 wait", "iss0 wait acc", "iss0 wait A"]
for i, n in enumerate(names):
    c = buf[:, 1 + i]
    print("%-16s min %8d  median %8d  max %8d" % (n, c.min(), np.median(c), c.max()))
print("clock: %.3f GHz" % (np.median(buf[:, 3] / buf[:, 1])))
