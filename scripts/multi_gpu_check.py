"""Run under torchrun (one rank per GPU): S-sharded NES attack vs the single-GPU run of the same attack (same Philox stream),
once with the one-shot peer-memory exchange (default) and once with the ncclAllReduce path (FB_NO_P2P semantics), then two
back-to-back sessions of different length.  Rank 0 prints the comparison."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fakebob_b200 import synth  # noqa: E402
from fakebob_b200.FAKEBOB import FakeBob  # noqa: E402
from fakebob_b200.gmm_ubm_OSI import gmm_OSI  # noqa: E402
from oracle import kaldi_feats as kf  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    holder = [None]
    if rank == 0:
        root = tempfile.mkdtemp(prefix="fb_mg_")
        tree = synth.build_gmm_tree(root, kf.voiced_features, n_speakers=3, C=256, n_ubm_utts=12, n_samples=32000,
                                    n_znorm_utts=4, em_iters=2)
        holder[0] = {k: tree[k] for k in ("pre_model_dir", "ubm", "models")}
        holder[0]["root"] = root
    dist.broadcast_object_list(holder, src=0)
    t = holder[0]
    model = gmm_OSI(os.path.join(tempfile.mkdtemp(), "g"), t["models"], t["ubm"], pre_model_dir=t["pre_model_dir"], device=local)
    audio = synth.synth_utterance(41, 1, 32000)
    thr = 1e3                                       # unreachable: every iteration runs (no early stop)
    hp = dict(max_iter=int(os.environ.get("MG_ITERS", "12")), samples_per_draw=int(os.environ.get("MG_S", "16")), plateau_length=3)
    single = None
    if rank == 0:
        fb1 = FakeBob("OSI", "untargeted", model, seed=123, verbose=False, **hp)
        fb1.attack(audio, None, threshold=thr)
        single = (fb1.final_adver.copy(), fb1.log.copy(), fb1.iters_done)
    dist.barrier()
    model._engine.comm_init_from_torch()
    print("rank", rank, "peer-memory exchange:", model._engine.p2p, flush=True)
    import time
    results = {}
    for mode in ("p2p", "nccl"):
        if mode == "nccl":
            os.environ["FB_NO_P2P"] = "1"           # fb_comm_p2p_attach checks it at every fb_nes_init
        fbm = FakeBob("OSI", "untargeted", model, seed=123, verbose=False, **hp)
        fbm.attack(audio, None, threshold=thr)
        dist.barrier()
        t0 = time.perf_counter()
        fbt = FakeBob("OSI", "untargeted", model, seed=124, verbose=False, max_iter=200, samples_per_draw=16)
        fbt.attack(audio, None, threshold=thr)
        dt = time.perf_counter() - t0
        advs = [None] * world
        dist.all_gather_object(advs, fbm.final_adver)
        results[mode] = (fbm.final_adver.copy(), fbm.log.copy())
        if rank == 0:
            same_across_ranks = all(np.array_equal(advs[0], a) for a in advs[1:])
            a1, l1, n1 = single
            print("[%s] world %d iters %d %d replicas identical: %s; 200-iteration attack %.1f it/s" % (mode, world, n1, fbm.iters_done, same_across_ranks, 200 / dt))
            print("[%s] adver agreement single vs sharded: %.6f" % (mode, np.mean(a1 == fbm.final_adver)))
            print("[%s] max |loss diff|: %.3e  max |final_loss diff|: %.3e" % (mode, np.abs(l1[:, 1] - fbm.log[:, 1]).max(), np.abs(l1[:, 2] - fbm.log[:, 2]).max()))
            assert same_across_ranks and n1 == fbm.iters_done
            assert np.mean(a1 == fbm.final_adver) > 0.999
    os.environ.pop("FB_NO_P2P", None)
    # a second, shorter utterance right after (new session number, other buffer sizes), still over peer memory
    audio2 = synth.synth_utterance(42, 2, 16000)
    fb2 = FakeBob("OSI", "untargeted", model, seed=7, verbose=False, max_iter=6, samples_per_draw=max(10, 2 * world + 2))   # every rank needs a pair
    fb2.attack(audio2, None, threshold=thr)
    advs = [None] * world
    dist.all_gather_object(advs, fb2.final_adver)
    if rank == 0:
        assert all(np.array_equal(advs[0], a) for a in advs[1:]) and fb2.iters_done == 6
        assert np.array_equal(results["p2p"][0], results["nccl"][0]) or np.mean(results["p2p"][0] == results["nccl"][0]) > 0.999
        print("p2p vs nccl adversarial audio identical fraction: %.6f" % np.mean(results["p2p"][0] == results["nccl"][0]))
        print("MULTI_GPU_OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
