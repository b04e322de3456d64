"""The i-vector / PLDA part of the oracle against independent implementations (scipy / sklearn / closed forms).
Like test_oracle_kaldi.py these do not pin the oracle to Kaldi (no Kaldi exists here); they guard it
against transcription errors."""
import numpy as np
import pytest
from scipy import stats
from scipy.linalg import cho_factor, cho_solve

from oracle.ivector import FullGmm, IvectorExtractor, PldaBackend, gaussian_selection, gselect_to_post

M_LOG_2PI = np.log(2 * np.pi)


def _full_gmm(C=12, D=6, seed=0):
    r = np.random.default_rng(seed)
    w = r.dirichlet(np.full(C, 3.0))
    mu = r.standard_normal((C, D)) * 2
    covs = []
    for _ in range(C):
        a = r.standard_normal((D, D)) * 0.3
        covs.append(a @ a.T + np.diag(r.uniform(0.5, 1.5, D)))
    covs = np.array(covs)
    inv = np.linalg.inv(covs)
    miv = np.einsum("cde,ce->cd", inv, mu)
    gc = np.array([np.log(w[c]) - 0.5 * (D * M_LOG_2PI + np.linalg.slogdet(covs[c])[1] + mu[c] @ inv[c] @ mu[c]) for c in range(C)])
    return FullGmm(w, miv, inv, gc), w, mu, covs


def test_full_covariance_loglikes_match_scipy():
    g, w, mu, covs = _full_gmm()
    r = np.random.default_rng(1)
    x = r.standard_normal(6)
    idx = np.array([0, 3, 7, 11])
    got = g.loglikes_preselect(x, idx)
    want = np.array([np.log(w[c]) + stats.multivariate_normal(mu[c], covs[c]).logpdf(x) for c in idx])
    assert np.abs(got - want).max() < 2e-4                       # float32 evaluation


def test_to_diag_keeps_means_and_marginal_variances():
    g, w, mu, covs = _full_gmm()
    d = g.to_diag()
    assert np.allclose(d.means(), mu, atol=1e-4)
    assert np.allclose(d.variances(), np.stack([np.diag(c) for c in covs]), rtol=1e-4)
    assert np.allclose(d.weights, w, atol=1e-7)


def test_gaussian_selection_and_pruned_posteriors():
    g, w, mu, covs = _full_gmm(C=40, D=6, seed=3)
    r = np.random.default_rng(4)
    X = (mu[r.integers(0, 40, 30)] + r.standard_normal((30, 6))).astype(np.float32)
    diag = g.to_diag()
    gsel, ll = gaussian_selection(diag, X, n=20)
    assert gsel.shape == (30, 20)
    for t in range(30):                                          # best first, exactly the 20 largest
        assert np.all(np.diff(ll[t, gsel[t]]) <= 0)
        assert set(gsel[t]) == set(np.argsort(-ll[t], kind="stable")[:20])
    post = gselect_to_post(g, X, gsel, min_post=0.025)
    for t in range(30):
        full_ll = np.array([np.log(w[c]) + stats.multivariate_normal(mu[c], covs[c]).logpdf(X[t].astype(np.float64)) for c in gsel[t]])
        p = np.exp(full_ll - full_ll.max())
        p /= p.sum()
        keep = p >= 0.025
        want = np.where(keep, p, 0.0)
        want /= want.sum()
        borderline = np.abs(p - 0.025) < 1e-4                    # a posterior within rounding of min_post may go either way
        if not borderline.any():
            assert np.abs(post[t] - want).max() < 2e-4
        assert abs(post[t].sum() - 1.0) < 1e-5 and np.all((post[t] == 0) | (post[t] >= 0.024))
    unpruned = gselect_to_post(g, X, gsel, min_post=0.0)
    assert np.all(unpruned > 0) and np.allclose(unpruned.sum(axis=1), 1.0, atol=1e-5)


def test_ivector_is_the_map_point_estimate():
    """w = argmax of the posterior of the total-variability model = solution of (I + sum_c gamma_c M_c' S_c^-1 M_c) w =
    sum_c M_c' S_c^-1 X_c + prior e_0; checked through a Cholesky solve and through the gradient of the objective."""
    r = np.random.default_rng(5)
    C, D, R = 10, 6, 7
    M = r.standard_normal((C, D, R)) * 0.3
    Sinv = np.stack([np.linalg.inv(a @ a.T + np.eye(D)) for a in r.standard_normal((C, D, D)) * 0.3])
    ie = IvectorExtractor(M, Sinv, prior_offset=100.0)
    gamma = r.uniform(0, 30, C)
    Xs = r.standard_normal((C, D)) * gamma[:, None]
    w = ie.extract(gamma, Xs)
    A = np.eye(R) + sum(gamma[c] * M[c].T @ Sinv[c] @ M[c] for c in range(C))
    b = sum(M[c].T @ Sinv[c] @ Xs[c] for c in range(C))
    b[0] += 100.0
    want = cho_solve(cho_factor(A), b)
    want[0] -= 100.0
    assert np.abs(w - want).max() < 1e-9
    # gradient of  -0.5 w'Aw + b'w  at the (un-offset) solution is zero
    wf = w.copy()
    wf[0] += 100.0
    assert np.abs(b - A @ wf).max() < 1e-8 * np.abs(b).max()
    # no data: the i-vector is the prior mean, i.e. zero after removing the offset
    assert np.abs(ie.extract(np.zeros(C), np.zeros((C, D)))).max() < 1e-12


def test_plda_llr_is_the_two_covariance_log_likelihood_ratio():
    r = np.random.default_rng(6)
    L = 5
    psi = np.sort(r.uniform(0.1, 4.0, L))[::-1]
    be = PldaBackend(np.zeros(8), np.eye(L, 8), np.zeros(L), np.eye(L), psi)
    u_train, u_test = r.standard_normal(L), r.standard_normal(L)
    for n in (1, 3):
        got = be.llr(u_train, u_test, n=n)
        # same speaker: u_test | mean of n enrolment vectors ~ N(n psi/(n psi+1) u_train, 1 + psi/(n psi+1)); different: N(0, 1+psi)
        same = stats.norm(n * psi / (n * psi + 1) * u_train, np.sqrt(1 + psi / (n * psi + 1))).logpdf(u_test).sum()
        diff = stats.norm(0.0, np.sqrt(1 + psi)).logpdf(u_test).sum()
        assert abs(got - (same - diff)) < 1e-10
    # joint-Gaussian check for n = 1: [u_train, u_test] ~ N(0, [[psi+1, psi],[psi, psi+1]]) vs independent
    joint = sum(stats.multivariate_normal([0, 0], [[p + 1, p], [p, p + 1]]).logpdf([a, b]) for p, a, b in zip(psi, u_train, u_test))
    indep = stats.norm(0, np.sqrt(1 + psi)).logpdf(u_train).sum() + stats.norm(0, np.sqrt(1 + psi)).logpdf(u_test).sum()
    assert abs(be.llr(u_train, u_test, n=1) - (joint - indep)) < 1e-9


def test_backend_length_normalisation():
    r = np.random.default_rng(7)
    R, L = 9, 4
    be = PldaBackend(r.standard_normal(R), r.standard_normal((L, R + 1)), r.standard_normal(L), np.eye(L) + 0.1 * r.standard_normal((L, L)),
                     np.array([3.0, 2.0, 1.0, 0.5]))
    v = be.prepare(r.standard_normal(R) * 5)
    assert v.dtype == np.float32 and abs(np.linalg.norm(v) - np.sqrt(L)) < 1e-4          # ivector-normalize-length: norm = sqrt(dim)
    u = be.transform(v)
    assert abs(np.dot(1.0 / (be.psi + 1.0), u * u) - L) < 1e-9                            # Plda::TransformIvector normalisation
    be0 = PldaBackend(np.zeros(R), np.eye(L, R), np.zeros(L), np.eye(L), be.psi)          # zero PLDA mean: scale invariant
    x = r.standard_normal(L)
    assert pytest.approx(be0.transform(2 * x)) == be0.transform(x)
