"""Generates tests/golden/nes_*.npz by running the REFERENCE's own FAKEBOB.py (imported from /root/reference,
unmodified) against the deterministic stub scorer, with the numpy global RNG seeded.  Run in the authoring
container only (the GPU box has no /root/reference):   python tests/golden/make_golden.py"""
import contextlib
import io
import os
import pickle
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")
import FAKEBOB as REF  # noqa: E402
from stub_scorer import CASES, StubScorer, make_audio  # noqa: E402


def main():
    for name, (task, attack_type, K, kw, hp) in CASES.items():
        model = StubScorer(K, 4000, sv=(task == "SV"))
        audio = make_audio(1)
        np.random.seed(2024)
        fb = REF.FakeBob(task, attack_type, model, **hp)
        cp = os.path.join(tempfile.mkdtemp(), "cp")
        with contextlib.redirect_stdout(io.StringIO()):
            adver, flag = fb.attack(audio.copy(), cp, **kw)
        with open(cp, "rb") as f:
            rows = pickle.load(f)
        dist = np.array([r[0] for r in rows], dtype=np.float64)
        loss = np.array([float(np.asarray(r[1]).reshape(-1)[0]) for r in rows], dtype=np.float64)
        out = os.path.join(HERE, "nes_%s.npz" % name)
        np.savez_compressed(out, adver=adver, flag=flag, distance=dist, adver_loss=loss, n_rows=len(rows))
        print(name, "flag", flag, "rows", len(rows), "->", out)
    # estimate_threshold golden (OSI)
    model = StubScorer(3, 4000, threshold=2.45)
    np.random.seed(7)
    fb = REF.FakeBob("OSI", "targeted", model, max_iter=10, samples_per_draw=10)
    with contextlib.redirect_stdout(io.StringIO()):
        res = fb.estimate_threshold(make_audio(2))
    np.savez_compressed(os.path.join(HERE, "nes_estimate_threshold.npz"), score=res[0], n_iters=res[1])
    print("estimate_threshold", res[:2])


if __name__ == "__main__":
    main()
