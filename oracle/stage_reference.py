"""TEST / BASELINE INFRASTRUCTURE: stage the reference's own Python sources where the GPU box can see them.

/root/reference exists only in the authoring container.  The reference is pure Python with no setup.py, so the "offline
install" of the bench contract (`pip install --target baseline/_ref /root/reference`) reduces to copying its top-level
modules: this script copies the two that are needed, UNMODIFIED, into baseline/_ref/ -- git-ignored (never part of the history), but not
gpurun-ignored, so it travels with the snapshot.  Users:
  * bench.py's CPU arm drives the reference's FAKEBOB.py (the NES loop itself) over the oracle scorers;
  * tests/test_gpu_attackmain.py executes the reference's attackMain.py against fakebob_b200/dropin.
Both skip gracefully when baseline/_ref is absent.  Run by __graft_entry__.build()."""
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["FAKEBOB.py", "attackMain.py"]          # the NES loop and the driver; the scorer modules are what this repo replaces


def stage():
    """-> list of staged files ([] when /root/reference is not present, e.g. on the GPU box)."""
    if not os.path.isdir(SRC):
        return []
    os.makedirs(DST, exist_ok=True)
    out = []
    for f in FILES:
        s = os.path.join(SRC, f)
        if os.path.exists(s):
            shutil.copyfile(s, os.path.join(DST, f))
            out.append(f)
    return out


if __name__ == "__main__":
    print("staged into %s: %s" % (DST, stage()))
