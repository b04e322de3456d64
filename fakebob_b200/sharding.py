"""How one NES draw is split across ranks (mirrors fb_nes_init in csrc/fb_nes.cu).

The S/2 antithetic pairs are block-partitioned: rank r owns pairs [floor(P r / W), floor(P (r+1) / W)); both
members of a pair stay on the same rank so one Philox draw serves both, and rank 0 additionally scores the
unperturbed audio (column 0 of FAKEBOB.py:234-237's batch).  Every rank writes its losses at their global
column index into a zero-initialised buffer [grad partial (N) | losses (S+1) | clean scores (K)], and ONE
all-reduce(sum) per iteration completes both the gradient estimate and the loss vector on every rank.
"""


def pair_range(pairs_total, rank, world):
    return (pairs_total * rank) // world, (pairs_total * (rank + 1)) // world


def global_column(pairs_total, pair, sign):
    """Column of `loss` / `noise` in the reference's layout [clean | +pairs | -pairs]."""
    return 1 + pair if sign > 0 else 1 + pairs_total + pair


def utterance_range(n_utterances, rank, world):
    """Independent attacks shard by utterance, no collective (BASELINE.json configs[4]: a batch of concurrent attack
    utterances over the GPUs of one box): rank r attacks utterances [floor(U r / W), floor(U (r+1) / W))."""
    return (n_utterances * rank) // world, (n_utterances * (rank + 1)) // world


def attack_many(make_attacker, audios, rank=0, world=1, **attack_kwargs):
    """Run one independent ``FakeBob.attack`` per utterance of this rank's shard.

    make_attacker() -> a FakeBob bound to this rank's device (built once, reused for every utterance);
    audios: sequence of 1-D float audios; attack_kwargs: forwarded to ``attack`` (threshold=, target=, ...; a callable value
    is called with the utterance index).  Returns {utterance index: (adver int16 (N,1), success flag)} for the shard; the
    caller gathers the dicts (e.g. ``torch.distributed.all_gather_object``) -- the attacks themselves never communicate."""
    lo, hi = utterance_range(len(audios), rank, world)
    fb = make_attacker()
    base_seed = fb.seed
    out = {}
    for u in range(lo, hi):
        # one Philox stream per utterance, a function of (seed, utterance index) only: the result of an attack does not
        # depend on how the utterances were sharded
        fb.seed = (base_seed + 0x9E3779B97F4A7C15 * (u + 1)) % (1 << 62)
        fb.draws = 0
        kw = {k: (v(u) if callable(v) else v) for k, v in attack_kwargs.items()}
        out[u] = fb.attack(audios[u], kw.pop("checkpoint_path", None), **kw)
    return out
