"""Run under torchrun (one rank per GPU): S-sharded NES attack with the NCCL gradient all-reduce vs the
single-GPU run of the same attack (same Philox stream).  Rank 0 prints the comparison."""
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
    hp = dict(max_iter=12, samples_per_draw=16, plateau_length=3)
    single = None
    if rank == 0:
        fb1 = FakeBob("OSI", "untargeted", model, seed=123, verbose=False, **hp)
        fb1.attack(audio, None, threshold=thr)
        single = (fb1.final_adver.copy(), fb1.log.copy(), fb1.iters_done)
    dist.barrier()
    model._engine.comm_init_from_torch()
    fbm = FakeBob("OSI", "untargeted", model, seed=123, verbose=False, **hp)
    fbm.attack(audio, None, threshold=thr)
    advs = [None] * world
    dist.all_gather_object(advs, fbm.final_adver)
    if rank == 0:
        same_across_ranks = all(np.array_equal(advs[0], a) for a in advs[1:])
        a1, l1, n1 = single
        print("world", world, "iters", n1, fbm.iters_done, "replicas identical:", same_across_ranks)
        print("adver agreement single vs sharded: %.6f" % np.mean(a1 == fbm.final_adver))
        print("max |loss diff|: %.3e  max |final_loss diff|: %.3e" % (np.abs(l1[:, 1] - fbm.log[:, 1]).max(), np.abs(l1[:, 2] - fbm.log[:, 2]).max()))
        assert same_across_ranks and n1 == fbm.iters_done
        assert np.mean(a1 == fbm.final_adver) > 0.999
        print("MULTI_GPU_OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
