"""Oracle (TEST INFRASTRUCTURE): Kaldi DiagGmm log-likelihood and MAP mean adaptation.
PARITY UNPINNED against real Kaldi (see oracle/__init__.py).

Reference call sites:
  gmm-global-get-frame-likes --average=true <model> <feats> ark,t:score  gmm_ubm_kaldiHelper.py:206-208
  gmm-global-acc-stats / gmm-global-est-map --update-flags=m             build_spk_models.py:197-219
                                                                         gmm-global-est-map.cc:81-83
Upstream algorithm: SURVEY.md Appendix A.7 (gmm/diag-gmm.cc DiagGmm::LogLikelihood,
matrix/kaldi-vector.cc LogSumExp, gmm/mle-diag-gmm.cc MapDiagGmmUpdate).
"""
import numpy as np

F32 = np.float32
K_MIN_LOG_DIFF_FLOAT = F32(np.log(np.finfo(np.float32).eps))   # -15.942385
M_LOG_2PI = 1.8378770664093454835606594728112


class DiagGmm:
    """Kaldi's stored form: weights (C), means_invvars (C,D), inv_vars (C,D), gconsts (C); all float32."""

    def __init__(self, weights, means_invvars, inv_vars, gconsts=None):
        self.weights = np.asarray(weights, dtype=F32)
        self.means_invvars = np.asarray(means_invvars, dtype=F32)
        self.inv_vars = np.asarray(inv_vars, dtype=F32)
        self.gconsts = np.asarray(gconsts, dtype=F32) if gconsts is not None else self.compute_gconsts()

    @classmethod
    def from_moments(cls, weights, means, variances):
        inv = (1.0 / np.asarray(variances, dtype=np.float64))
        return cls(weights, (np.asarray(means, dtype=np.float64) * inv).astype(F32), inv.astype(F32))

    @property
    def num_gauss(self):
        return self.weights.shape[0]

    @property
    def dim(self):
        return self.inv_vars.shape[1]

    def means(self):
        return (self.means_invvars.astype(np.float64) / self.inv_vars.astype(np.float64))

    def variances(self):
        return 1.0 / self.inv_vars.astype(np.float64)

    def compute_gconsts(self):
        """DiagGmm::ComputeGconsts (double accumulation, stored float)."""
        iv = self.inv_vars.astype(np.float64)
        miv = self.means_invvars.astype(np.float64)
        D = iv.shape[1]
        gc = np.log(self.weights.astype(np.float64)) - 0.5 * D * M_LOG_2PI
        gc = gc + (0.5 * np.log(iv) - 0.5 * miv * miv / iv).sum(axis=1)
        return gc.astype(F32)

    def loglikes(self, X):
        """Per-frame per-component log-likelihoods, float32 (two sgemv per frame in Kaldi)."""
        X = np.asarray(X, dtype=F32)
        ll = X @ self.means_invvars.T
        ll += (X * X) @ (F32(-0.5) * self.inv_vars).T
        ll += self.gconsts[None, :]
        return ll.astype(F32)

    def frame_loglikes(self, X):
        """DiagGmm::LogLikelihood per row: Kaldi LogSumExp (float exp, double sum, float cutoff)."""
        return log_sum_exp_rows(self.loglikes(X))

    def avg_loglike(self, X):
        """gmm-global-get-frame-likes --average=true: Sum (double) / T, float."""
        fl = self.frame_loglikes(X)
        if fl.shape[0] == 0:
            raise ValueError("no voiced frames")
        return F32(F32(fl.astype(np.float64).sum()) / F32(fl.shape[0]))

    def posteriors(self, X):
        ll = self.loglikes(X).astype(np.float64)
        ll -= ll.max(axis=1, keepdims=True)
        p = np.exp(ll)
        return p / p.sum(axis=1, keepdims=True)

    def component_posteriors(self, X):
        """DiagGmm::ComponentPosteriors (gmm/diag-gmm.cc): float log-likelihoods, then VectorBase<float>::ApplySoftMax
        (float max, float exp, float running sum in index order, one float scale)."""
        ll = self.loglikes(X)
        mx = ll.max(axis=1, keepdims=True)
        e = np.exp((ll - mx).astype(F32)).astype(F32)
        s = np.cumsum(e, axis=1, dtype=F32)[:, -1:]
        return (e * (F32(1.0) / s).astype(F32)).astype(F32)

    def map_adapt_means(self, X, tau=10.0):
        """gmm-global-acc-stats --update-flags=m + gmm-global-est-map --update-flags=m (build_spk_models.py:202-219; the
        reference's gmm-global-est-map.cc calls MapDiagGmmUpdate, mean_tau = 10): AccumDiagGmm::AccumulateFromPosteriors
        adds post_d[c] and post_d[c] * x_d per frame in double, frame order; MapDiagGmmUpdate sets
        mean = (mean_acc + tau * old_mean) / (occ + tau); weights and variances stay; ComputeGconsts."""
        post = self.component_posteriors(X).astype(np.float64)            # (T, C)
        X64 = np.asarray(np.asarray(X, dtype=F32), dtype=np.float64)
        C, D = self.means_invvars.shape
        occ = np.zeros(C)
        mean_acc = np.zeros((C, D))
        for t in range(post.shape[0]):
            occ += post[t]
            mean_acc += post[t][:, None] * X64[t][None, :]
        old = self.means()
        new = (mean_acc + tau * old) * (1.0 / (occ + tau))[:, None]
        iv = self.inv_vars.astype(np.float64)
        out = DiagGmm(self.weights, (new * iv).astype(F32), self.inv_vars)
        out.occupancy = occ
        return out


def log_sum_exp_rows(ll):
    """VectorBase<float>::LogSumExp(prune=-1) applied to each row of a float32 matrix."""
    ll = np.asarray(ll, dtype=F32)
    m = ll.max(axis=1)
    cutoff = (m + K_MIN_LOG_DIFF_FLOAT).astype(F32)
    e = np.exp((ll - m[:, None]).astype(F32)).astype(F32)
    e = np.where(ll >= cutoff[:, None], e, F32(0.0))
    s = e.astype(np.float64).sum(axis=1)
    return (m.astype(np.float64) + np.log(s)).astype(F32)
