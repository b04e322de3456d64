"""Short run of the bench workload for ncu: a few NES iterations of config C2 without CUDA-graph replay.
Usage (under gpurun): FB_NO_GRAPH=1 ncu --profile-from-start off ... python scripts/profile_iter.py [n_iters]"""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from fakebob_b200 import synth  # noqa: E402
from fakebob_b200.FAKEBOB import FakeBob  # noqa: E402
from fakebob_b200.gmm_ubm_OSI import gmm_OSI  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    root = tempfile.mkdtemp(prefix="fakebob_prof_")
    tree = bench.build_workload_gpu(root, 0)
    model = gmm_OSI(os.path.join(root, "grp"), tree["models"], tree["ubm"], pre_model_dir=tree["pre_model_dir"], device=0)
    audio = synth.synth_utterance(0, 0, bench.N_SAMPLES)
    fb = FakeBob("OSI", "untargeted", model, max_iter=n, samples_per_draw=bench.S_DRAW, seed=1, verbose=False, iters_per_launch=n)
    # ncu --profile-from-start off: only the NES iterations are inside the profiled range (the workload builder above
    # launches the same kernels on single utterances)
    import ctypes
    rt = ctypes.CDLL("libcudart.so")
    rt.cudaProfilerStart()
    fb.attack(audio, None, threshold=1e3)
    rt.cudaProfilerStop()
    print("iterations:", fb.iters_done, "voiced rows:", model._engine.voiced_rows(), "launches:", model._engine.kernel_launches())


if __name__ == "__main__":
    main()
