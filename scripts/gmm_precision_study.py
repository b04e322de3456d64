"""Operand-rounding error of GMM contraction variants (numpy emulation, float64 accumulation).

Variants of  ll_m = g_m + x.(mu_m/var) - 0.5 x^2.(1/var)  with fp16 hi/lo split operands:
  full3   : every model, 3 terms hi.hi + lo.hi + hi.lo                      (round-1 kernel)
  delta1  : UBM with 3 terms; speaker m = UBM + x_hi . dW_hi                (dW = (mu_m - mu_ubm)/var, dg = g_m - g_ubm)
  delta2a : ... + x_lo . dW_hi
  delta2b : ... + x_hi . dW_lo
  delta3  : all three delta terms
Prints per-frame LL error, per-utterance average LL error and score (LL_spk - LL_ubm) error vs float64.
"""
import os, sys, time, pickle
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fakebob_b200 import synth, kaldi_io
from oracle import kaldi_feats as kf

def h16(a):
    return a.astype(np.float16).astype(np.float64)

def split(a):
    hi = h16(a)
    lo = h16(a - hi)
    return hi, lo

def lse(ll):
    m = ll.max(axis=1, keepdims=True)
    return (m + np.log(np.exp(ll - m).sum(axis=1, keepdims=True)))[:, 0]

def main():
    root = "/tmp/c2"
    cache = os.path.join(root, "tree.pkl")
    if os.path.exists(cache):
        tree = pickle.load(open(cache, "rb"))
    else:
        t0 = time.time()
        tree = synth.build_gmm_tree(root, kf.voiced_features, n_speakers=5, C=2048, n_ubm_utts=64, n_samples=80000)
        print("tree built in %.1f s" % (time.time() - t0))
        pickle.dump(tree, open(cache, "wb"))
    gm = [kaldi_io.read_diag_gmm(tree["ubm"])] + [kaldi_io.read_diag_gmm(m[2]) for m in tree["models"]]
    iv = gm[0]["inv_vars"].astype(np.float64)
    # feature scale as fb_finalize_gmms
    mu0 = gm[0]["means_invvars"].astype(np.float64) / iv
    w0 = gm[0]["weights"].astype(np.float64)
    rms = np.sqrt((w0[:, None] * (mu0 * mu0 + 1.0 / iv)).sum(0) / w0.sum())
    sc = 2.0 ** (-np.rint(np.log2(rms)))
    log2e = 1.4426950408889634
    W1 = [g["means_invvars"].astype(np.float64) / sc * log2e for g in gm]          # (C, D)
    W2 = -0.5 * iv / (sc * sc) * log2e
    G = [g["gconsts"].astype(np.float64) * log2e for g in gm]
    res = {}
    for utt in range(3):
        a = synth.to_int16(synth.synth_utterance(utt, 0, 80000))
        X = kf.voiced_features(a).astype(np.float64)
        xs = X * sc
        x2 = (xs.astype(np.float32) * xs.astype(np.float32)).astype(np.float64)    # kernel squares in float
        xh, xl = split(xs)
        qh, ql = split(x2)
        exact = [xs @ W1[m].T + x2 @ W2.T + G[m] for m in range(6)]               # log2 domain
        ex_ll = [lse(e * np.log(2)) for e in exact]
        w2h, w2l = split(W2)
        Q = qh @ w2h.T + ql @ w2h.T + qh @ w2l.T
        def t3(Wm):
            wh, wl = split(Wm)
            return xh @ wh.T + xl @ wh.T + xh @ wl.T
        def g3(g):
            g0 = h16(g); g1 = h16(g - g0); g2 = h16(g - g0 - g1)
            return g0 + g1 + g2
        T0 = t3(W1[0]) + g3(G[0])
        variants = {}
        variants["full3"] = [Q + t3(W1[m]) + g3(G[m]) for m in range(6)]
        mus = mu0 * sc                                                           # UBM means in scaled units
        for name in ("delta1", "delta2a", "delta2b", "delta3", "delta1f", "delta2af"):
            outs = [Q + T0]
            for m in range(1, 6):
                dW = W1[m] - W1[0]
                dg = G[m] - G[0]
                dh, dl = split(dW)
                d = xh @ dh.T
                if name in ("delta2a", "delta3", "delta2af"):
                    d = d + xl @ dh.T
                if name in ("delta2b", "delta3"):
                    d = d + xh @ dl.T
                if name.endswith("f"):                                             # fold E[x . dl] at x = mu_c into the constant
                    dg = dg + ((dW - dh) * mus).sum(1)
                outs.append(Q + T0 + d + g3(dg))
            variants[name] = outs
        for name, outs in variants.items():
            ll = [lse(o * np.log(2)) for o in outs]
            fe = max(np.abs(ll[m] - ex_ll[m]).max() for m in range(6))
            fe_rms = np.sqrt(np.mean([(ll[m] - ex_ll[m]) ** 2 for m in range(6)]))
            ae = max(abs(ll[m].mean() - ex_ll[m].mean()) for m in range(6))
            se = max(abs((ll[m].mean() - ll[0].mean()) - (ex_ll[m].mean() - ex_ll[0].mean())) for m in range(1, 6))
            res.setdefault(name, []).append((fe, fe_rms, ae, se))
        # float32 Kaldi-style reference error for scale: sgemv in float32
        f32 = [(X.astype(np.float32) @ gm[m]["means_invvars"].T + (X.astype(np.float32) ** 2) @ (-0.5 * gm[m]["inv_vars"]).T + gm[m]["gconsts"]).astype(np.float64) for m in range(6)]
        ll = [lse(o) for o in f32]
        fe = max(np.abs(ll[m] - ex_ll[m]).max() for m in range(6))
        fe_rms = np.sqrt(np.mean([(ll[m] - ex_ll[m]) ** 2 for m in range(6)]))
        ae = max(abs(ll[m].mean() - ex_ll[m].mean()) for m in range(6))
        se = max(abs((ll[m].mean() - ll[0].mean()) - (ex_ll[m].mean() - ex_ll[0].mean())) for m in range(1, 6))
        res.setdefault("float32_sgemm", []).append((fe, fe_rms, ae, se))
        print("utt %d rows %d  score(ex) %s" % (utt, X.shape[0], np.round([ex_ll[m].mean() - ex_ll[0].mean() for m in range(1, 6)], 4)))
    print("%-14s %12s %12s %12s %12s" % ("variant", "frameLL max", "frameLL rms", "avgLL max", "score max"))
    for name, v in res.items():
        v = np.array(v)
        print("%-14s %12.3e %12.3e %12.3e %12.3e" % (name, v[:, 0].max(), v[:, 1].max(), v[:, 2].max(), v[:, 3].max()))

if __name__ == "__main__":
    main()
