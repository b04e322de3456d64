"""Short run of a bench workload for ncu / compute-sanitizer: a few NES iterations of config C2 (default) or C3 without
CUDA-graph replay, inside a cudaProfilerStart/Stop range.
Usage (under gpurun): FB_NO_GRAPH=1 ncu --profile-from-start off ... python scripts/profile_iter.py [n_iters] [C2|C3|C4]"""
import ctypes
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from fakebob_b200 import synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    name = sys.argv[2] if len(sys.argv) > 2 else "C2"
    c = bench.CONFIGS[name]
    root = tempfile.mkdtemp(prefix="fakebob_prof_")
    tree = bench.build_gmm_tree_gpu(root, 0, c["K"] if c["arch"] == "gmm" else 1)
    iv_models = bench.build_iv_speakers_gpu(root, tree, 0, c["K"])["models"] if c["arch"] == "iv" else None
    model = bench.make_model(name, tree, iv_models, 0, os.path.join(root, "grp"))
    audio = synth.synth_utterance(0, 0, bench.N_SAMPLES)
    fb = bench.make_attacker(name, model, n, 1)
    fb.iters_per_launch = n
    fb.attack(audio, None, **bench.attack_kwargs(name))             # warm-up: allocations, module load
    fb = bench.make_attacker(name, model, n, 2)
    fb.iters_per_launch = n
    # only the NES iterations are inside the profiled range (the workload builder launches the same kernels on single utterances)
    rt = ctypes.CDLL("libcudart.so")
    rt.cudaProfilerStart()
    fb.attack(audio, None, **bench.attack_kwargs(name))
    rt.cudaProfilerStop()
    print("config", name, "iterations:", fb.iters_done, "voiced rows:", model._engine.voiced_rows(), "launches:", model._engine.kernel_launches())


if __name__ == "__main__":
    main()
