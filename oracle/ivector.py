"""Oracle (TEST INFRASTRUCTURE): i-vector extraction and PLDA scoring as the reference runs them through Kaldi.
PARITY UNPINNED against real Kaldi (see oracle/__init__.py).

Reference call sites:
  sid/extract_ivectors.sh ........ ivector_PLDA_kaldiHelper.py:202-211
      gmm-gselect --n=20 "fgmm-global-to-gmm final.ubm -|" | fgmm-global-gselect-to-post --min-post=0.025 final.ubm
      | scale-post 1.0 | ivector-extract final.ie
  ivector-plda-scoring ........... ivector_PLDA_kaldiHelper.py:262-271
      "ivector-copy-plda --smoothing=0.0 plda - |"  with both sides piped through
      ivector-subtract-global-mean mean.vec | transform-vec transform.mat | ivector-normalize-length
Upstream algorithm: SURVEY.md Appendix A.8 (gmm/full-gmm.cc, gmm/diag-gmm.cc GaussianSelection,
fgmmbin/fgmm-global-gselect-to-post.cc, ivector/ivector-extractor.cc, ivector/plda.cc).

float32 where Kaldi's BaseFloat is float (selection, full-covariance log-likelihoods, posteriors, the stored
i-vector and the LDA / length-norm chain), float64 where Kaldi uses double (statistics, extractor, PLDA).
"""
import numpy as np

from . import kaldi_feats as kf
from .diag_gmm import DiagGmm

F32 = np.float32
M_LOG_2PI = 1.8378770664093454835606594728112


class FullGmm:
    def __init__(self, weights, means_invcovars, inv_covars, gconsts):
        self.weights = np.asarray(weights, dtype=F32)
        self.means_invcovars = np.asarray(means_invcovars, dtype=F32)
        self.inv_covars = np.asarray(inv_covars, dtype=F32)          # (C, D, D) symmetric
        self.gconsts = np.asarray(gconsts, dtype=F32)

    def to_diag(self):
        """fgmm-global-to-gmm (DiagGmm::CopyFromFullGmm): same weights and means, variances = diag(covariance)."""
        C, D = self.means_invcovars.shape
        means = np.empty((C, D))
        dvar = np.empty((C, D))
        for c in range(C):
            cov = np.linalg.inv(self.inv_covars[c].astype(np.float64))
            means[c] = cov @ self.means_invcovars[c].astype(np.float64)
            dvar[c] = np.diag(cov)
        inv_vars = (1.0 / dvar.astype(F32)).astype(F32)               # stored float
        means_f = means.astype(F32)
        return DiagGmm(self.weights, (means_f * inv_vars).astype(F32), inv_vars)

    def loglikes_preselect(self, x, idx):
        """FullGmm::LogLikelihoodsPreselect: gconst + (S^-1 mu).x - 0.5 x'S^-1 x for the selected components (float)."""
        x = np.asarray(x, dtype=F32)
        S = self.inv_covars[idx]                                       # (n, D, D)
        quad = np.einsum("d,nde,e->n", x, S, x, dtype=F32)
        lin = self.means_invcovars[idx] @ x
        return (self.gconsts[idx] + lin - F32(0.5) * quad).astype(F32)


def gaussian_selection(diag, X, n=20):
    """gmm-gselect --n=20: indices of the n best components per frame, best first."""
    ll = diag.loglikes(X)
    idx = np.argsort(-ll, axis=1, kind="stable")[:, :n]
    return idx, ll


def gselect_to_post(full, X, gselect, min_post=0.025):
    """fgmm-global-gselect-to-post --min-post=0.025 -> (T, n) posteriors aligned with gselect (0 = pruned)."""
    T, n = gselect.shape
    post = np.zeros((T, n), dtype=F32)
    for t in range(T):
        ll = full.loglikes_preselect(X[t], gselect[t])
        m = ll.max()
        e = np.exp((ll - m).astype(F32)).astype(F32)
        ssum = F32(0.0)
        for v in e:                                                   # ApplySoftMax: sequential float sum, Scale(1/sum)
            ssum = F32(ssum + v)
        p = (e * F32(1.0 / float(ssum))).astype(F32)
        if min_post != 0.0:
            mx = int(np.argmax(p))
            p = np.where(p < F32(min_post), F32(0.0), p)
            s = F32(p.astype(np.float64).sum())
            if s == 0.0:
                p[mx] = 1.0
            else:
                p = (p * F32(1.0 / float(s))).astype(F32)
        post[t] = p
    return post


class IvectorExtractor:
    """ivector/ivector-extractor.cc without weight projection (w_ empty), max-count 0."""

    def __init__(self, M, sigma_inv, prior_offset):
        self.M = np.asarray(M, dtype=np.float64)                       # (C, D, R)
        self.sigma_inv = np.asarray(sigma_inv, dtype=np.float64)       # (C, D, D)
        self.prior_offset = float(prior_offset)
        # ComputeDerivedVars: Sigma_inv_M_[c] = Sigma_inv_[c] * M_[c];  U_[c] = M_[c]' Sigma_inv_[c] M_[c]
        self.sigma_inv_M = np.einsum("cde,cer->cdr", self.sigma_inv, self.M)
        self.U = np.einsum("cdr,cds->crs", self.M, self.sigma_inv_M)   # (C, R, R)

    @property
    def ivector_dim(self):
        return self.M.shape[2]

    def extract(self, gamma, Xs):
        """gamma (C) float64, Xs (C, D) float64 first-order stats -> i-vector (R) float64 (prior offset removed)."""
        lin = np.einsum("cdr,cd->r", self.sigma_inv_M, Xs)
        quad = np.einsum("c,crs->rs", gamma, self.U)
        lin[0] += self.prior_offset
        quad = quad + np.eye(self.ivector_dim)
        w = np.linalg.solve(quad, lin)
        w[0] -= self.prior_offset
        return w


class PldaBackend:
    """mean.vec / transform.mat / plda as used by ivector-plda-scoring --normalize-length=true (ivector/plda.cc)."""

    def __init__(self, mean_vec, transform_mat, plda_mean, plda_transform, plda_psi):
        self.mean_vec = np.asarray(mean_vec, dtype=F32)
        self.transform_mat = np.asarray(transform_mat, dtype=F32)      # (L, R) or (L, R+1) with offset column
        self.plda_mean = np.asarray(plda_mean, dtype=np.float64)
        self.plda_transform = np.asarray(plda_transform, dtype=np.float64)
        self.psi = np.asarray(plda_psi, dtype=np.float64)
        self.offset = -self.plda_transform @ self.plda_mean            # Plda::ComputeDerivedVars

    def prepare(self, ivector):
        """ivector-subtract-global-mean | transform-vec | ivector-normalize-length   (float32 chain)."""
        v = (np.asarray(ivector, dtype=F32) - self.mean_vec).astype(F32)
        T = self.transform_mat
        R = v.shape[0]
        if T.shape[1] == R + 1:
            out = (T[:, R] + T[:, :R] @ v).astype(F32)
        else:
            out = (T @ v).astype(F32)
        norm = F32(np.sqrt(np.sum(out.astype(np.float64) ** 2)))
        ratio = F32(norm / F32(np.sqrt(F32(out.shape[0]))))
        return (out * (F32(1.0) / ratio)).astype(F32)

    def transform(self, v, num_examples=1):
        """Plda::TransformIvector with normalize_length=true, simple_length_norm=false (double)."""
        u = self.offset + self.plda_transform @ np.asarray(v, dtype=np.float64)
        inv_covar = 1.0 / (self.psi + 1.0 / num_examples)
        factor = np.sqrt(u.shape[0] / np.dot(inv_covar, u * u))
        return u * factor

    def llr(self, u_train, u_test, n=1):
        """Plda::LogLikelihoodRatio (double)."""
        psi = self.psi
        L = psi.shape[0]
        mean = n * psi / (n * psi + 1.0) * u_train
        var = 1.0 + psi / (n * psi + 1.0)
        given = -0.5 * (np.log(var).sum() + M_LOG_2PI * L + np.dot((u_test - mean) ** 2, 1.0 / var))
        var0 = 1.0 + psi
        without = -0.5 * (np.log(var0).sum() + M_LOG_2PI * L + np.dot(u_test ** 2, 1.0 / var0))
        return given - without


class IvectorSystem:
    def __init__(self, full_ubm, extractor, backend, cfg=None, num_gselect=20, min_post=0.025):
        self.full = full_ubm
        self.diag = full_ubm.to_diag()
        self.extractor = extractor
        self.backend = backend
        self.cfg = cfg or kf.FeatConfig()
        self.num_gselect = num_gselect
        self.min_post = min_post

    def posteriors(self, X):
        gsel, _ = gaussian_selection(self.diag, X, self.num_gselect)
        return gsel, gselect_to_post(self.full, X, gsel, self.min_post)

    def stats(self, X, gsel, post):
        C, D = self.full.means_invcovars.shape
        gamma = np.zeros(C)
        Xs = np.zeros((C, D))
        Xd = np.asarray(X, dtype=np.float64)
        for t in range(X.shape[0]):
            for j in range(gsel.shape[1]):
                p = float(post[t, j])
                if p != 0.0:
                    gamma[gsel[t, j]] += p
                    Xs[gsel[t, j]] += p * Xd[t]
        return gamma, Xs

    def extract_from_features(self, X):
        gsel, post = self.posteriors(X)
        gamma, Xs = self.stats(X, gsel, post)
        return self.extractor.extract(gamma, Xs).astype(F32)          # ivector-extract writes Vector<BaseFloat>

    def extract(self, wave_int16):
        X = kf.voiced_features(wave_int16, self.cfg)
        if X.shape[0] == 0:
            raise ValueError("no voiced frames")
        return self.extract_from_features(X)

    def plda_scores(self, enrolled, test):
        """enrolled (K,R), test (B,R) raw i-vectors -> (B,K) float64 LLR (trials = enrolled x test, n=1 each)."""
        ue = [self.backend.transform(self.backend.prepare(e)) for e in np.atleast_2d(enrolled)]
        ut = [self.backend.transform(self.backend.prepare(t)) for t in np.atleast_2d(test)]
        return np.array([[self.backend.llr(e, t) for e in ue] for t in ut], dtype=np.float64)


def load_system(pre_model_dir, cfg=None):
    """Build the oracle system from a Kaldi-format pre-models/ tree, parsed with the oracle's own reader (oracle/kaldi_files.py)."""
    import os
    from . import kaldi_files as kaldi_io
    fg = kaldi_io.read_full_gmm(os.path.join(pre_model_dir, "final.ubm"))
    ie = kaldi_io.read_ivector_extractor(os.path.join(pre_model_dir, "final.ie"))
    pl = kaldi_io.read_plda(os.path.join(pre_model_dir, "plda"))
    full = FullGmm(fg["weights"], fg["means_invcovars"], fg["inv_covars"], fg["gconsts"])
    backend = PldaBackend(kaldi_io.read_vector(os.path.join(pre_model_dir, "mean.vec")),
                          kaldi_io.read_matrix(os.path.join(pre_model_dir, "transform.mat")), pl["mean"], pl["transform"], pl["psi"])
    return IvectorSystem(full, IvectorExtractor(ie["M"], ie["sigma_inv"], ie["prior_offset"]), backend, cfg)
