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
