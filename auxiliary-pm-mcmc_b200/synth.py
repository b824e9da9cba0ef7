"""Synthetic GP-probit data sets of the UCI pima / breast shapes (the UCI files are not available
offline).  Recipe: SURVEY.md §8(d).  Host-side, one-off data generation -- not on the hot path.

Shapes: 'pima' (768, 8), 'breast' (682, 9), 'large' (8192, 16).  Priors / proposal scales / slice
widths are those of the reference notebooks (experiment_notebooks/*.ipynb cell 4 and cell 8).
"""
import numpy as np

SHAPES = {'pima': (768, 8), 'breast': (682, 9), 'large': (8192, 16)}


def _ard_cov(X, theta, eps):
    Xs = X / np.exp(theta[1:])[None]
    sq = (Xs**2).sum(1)
    d2 = np.maximum(sq[:, None] + sq[None, :] - 2. * Xs.dot(Xs.T), 0.)
    K = np.exp(theta[0]) * np.exp(-0.5 * d2)
    K[np.diag_indices(X.shape[0])] = np.exp(theta[0]) + eps
    return K


def make_dataset(n, D, seed=0, eps=1e-8):
    """Returns (X, y, theta_true): X normalised to zero mean / unit sd per column
    (gpdemo/utils.py:86-105), y in {-1,+1} drawn from a probit GP with an ARD kernel."""
    rs = np.random.RandomState(seed)
    X = rs.normal(size=(n, D))
    X = (X - X.mean(0)[None]) / X.std(0)[None]
    theta_true = np.r_[0.5, np.full(D, np.log(2. * np.sqrt(D / 8.)))]
    K = _ard_cov(X, theta_true, eps)
    f = np.linalg.cholesky(K).dot(rs.normal(size=n))
    from math import erf, sqrt
    Phi = 0.5 * (1. + np.array([erf(v / sqrt(2.)) for v in f]))
    y = np.where(rs.uniform(size=n) < Phi, 1., -1.)
    return np.ascontiguousarray(X), y, theta_true


def named_dataset(name, seed=0):
    n, D = SHAPES[name]
    return make_dataset(n, D, seed)


def prior_params(D):
    """Log-Gamma prior hyper-parameters of the notebooks (nb cell 8)."""
    return dict(a_sigma=1.1, b_sigma=0.1, a_tau=1., b_tau=1. / D**0.5)


def draw_theta_prior(rs, D, ard=True):
    """theta_init as the notebooks draw it (nb cell 14): log of Gamma draws; for the ARD kernel the
    same tau prior is used for every input dimension (our extension)."""
    p = prior_params(D)
    th = [np.log(rs.gamma(p['a_sigma'], 1. / p['b_sigma']))]
    for _ in range(D if ard else 1):
        th.append(np.log(rs.gamma(p['a_tau'], 1. / p['b_tau'])))
    return np.array(th)


def bulk_thetas(n_chain, D, seed=1234, ard=True, spread=0.3):
    """Kernel parameters from the bulk of the posterior region (moderate cond(K)): theta_true plus a
    Gaussian perturbation.  Used by the benchmark so every chain does a typical Newton solve."""
    rs = np.random.RandomState(seed)
    base = np.r_[0.5, np.full(D if ard else 1, np.log(2. * np.sqrt(D / 8.)))]
    return base[None] + spread * rs.normal(size=(n_chain, base.shape[0]))
