"""CPU oracle: numpy/scipy restatement of the reference's pseudo-marginal likelihood hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under `auxiliary-pm-mcmc_b200/` (the product) may import this
module; only tests/, `__graft_entry__.smoke()` and bench.py's cpu_baseline / `--impl reference`
legs use it, and there only as the checker / CPU baseline -- never as the thing shipped.

Parity status: the reference (matt-graham/auxiliary-pm-mcmc) has NO tests, fixtures or golden vectors
for this path ("parity unpinned by the reference's own tests").  The oracle is
instead pinned against OUTPUTS OF THE REFERENCE ITSELF, executed unmodified in the build container
through oracle/ref_loader.py: tests/golden/*.npz are produced by oracle/gen_golden.py from the real
reference and tests/test_oracle_vs_golden.py checks this restatement against them (bit-for-bit for the
covariance builders, <= 1e-12 relative for everything that goes through LAPACK).

Third-party arithmetic the reference leans on (not vendored; README.md:19-20 pins numpy 1.9.2 /
scipy 0.16.0 in prose only): LAPACK dpotrf/dpotrs, BLAS dgemm/dgemv, scipy.special.log_ndtr,
scipy logsumexp.  Here they are this container's numpy 2.3 / scipy 1.18 (OpenBLAS 0.3.30).

Each function cites the reference file:line it restates (paths relative to /root/reference).
"""
import ctypes as ct
import os
import subprocess

import numpy as np
import scipy.linalg as la
from scipy.special import log_ndtr, logsumexp, gammaln

HALF_LOG_2PI = 0.5 * np.log(2 * np.pi)


class MaximumIterationsExceededError(Exception):
    """gpdemo/latent_posterior_approximations.py:17-19"""


class InvalidCovarianceMatrixError(Exception):
    """gpdemo/estimators.py:85-87"""


# ----------------------------------------------------------------------------------------------
# covariance builders  (gpdemo/kernels.pyx)
# ----------------------------------------------------------------------------------------------

_HERE = os.path.dirname(os.path.abspath(__file__))
_clib = None


def _kernels_clib():
    """Plain-C restatement (oracle/kernels_oracle.c), compiled on first use with gcc."""
    global _clib
    if _clib is None:
        so = os.path.join(_HERE, 'libkernels_oracle.so')
        src = os.path.join(_HERE, 'kernels_oracle.c')
        if not os.path.isfile(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-fPIC', '-shared', src, '-o', so, '-lm'])
        _clib = ct.CDLL(so)
        for fn in (_clib.oracle_iso_se_kernel, _clib.oracle_ard_se_kernel):
            fn.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_double, ct.c_int, ct.c_int]
            fn.restype = None
    return _clib


def _call_kernel(fn, K, X, theta, epsilon, n_theta):
    X = np.ascontiguousarray(X, dtype=np.float64)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    n, D = X.shape
    assert theta.shape == (n_theta(D),) and K.shape == (n, n) and K.dtype == np.float64
    out = K if K.flags.c_contiguous else np.empty((n, n))
    fn(out.ctypes.data, X.ctypes.data, theta.ctypes.data, float(epsilon), n, D)
    if out is not K:
        K[:, :] = out


def isotropic_squared_exponential_kernel(K, X, theta, epsilon=1e-8):
    """gpdemo/kernels.pyx:12-49, in place.  Scalar C restatement (oracle/kernels_oracle.c): libm exp and
    the reference's operation order, so the result is bit-identical to the Cython module."""
    _call_kernel(_kernels_clib().oracle_iso_se_kernel, K, X, theta, epsilon, lambda D: 2)
    return None


def diagonal_squared_exponential_kernel(K, X, theta, epsilon=1e-8):
    """gpdemo/kernels.pyx:52-90 (ARD), in place; see isotropic_squared_exponential_kernel."""
    _call_kernel(_kernels_clib().oracle_ard_se_kernel, K, X, theta, epsilon, lambda D: D + 1)
    return None


# ----------------------------------------------------------------------------------------------
# Expectation propagation -- NO REFERENCE COUNTERPART (the project brief names EP, the reference has only the Laplace
# approximation: SURVEY.md App. D).  Two restatements of Rasmussen & Williams, GPML (2006), Alg. 3.5 for the probit
# likelihood: the textbook sequential sweep (rank-one updates of Sigma) and the variant the CUDA path runs, in which
# all sites are updated from the same (mu, diag Sigma) before Sigma and mu are recomputed ("parallel EP").  Both stop
# at the same fixed point; tests/test_oracle_vs_golden.py pins the parallel one to the textbook one.
# ----------------------------------------------------------------------------------------------

def _probit_moments(y, mu_c, s2_c):
    """Mean and variance of the tilted distribution Phi(y f) N(f | mu_c, s2_c)  (GPML eq. 3.58)."""
    den = np.sqrt(1. + s2_c)
    z = y * mu_c / den
    r = np.exp(-0.5 * z * z - log_ndtr(z) - HALF_LOG_2PI)          # N(z) / Phi(z)
    mu_h = mu_c + y * s2_c * r / den
    s2_h = s2_c - s2_c * s2_c * r * (z + r) / (1. + s2_c)
    return mu_h, s2_h


def _ep_posterior(K, nu, tau):
    """Sigma = K - K S^1/2 B^-1 S^1/2 K, mu = Sigma nu~ (GPML eqs. 3.53, 3.68) in the arithmetic the CUDA path uses:
    a = nu~ - S^1/2 B^-1 (S^1/2 K nu~), mu = K a; Z = K S^1/2 L^-T, diag(Sigma) = diag(K) - rowsum(Z^2)."""
    n = K.shape[0]
    Ss = np.sqrt(tau)
    B = np.eye(n) + (Ss[:, None] * K) * Ss[None, :]
    L = la.cholesky(B, lower=True)
    t = Ss * K.dot(nu)
    a = nu - Ss * la.cho_solve((L, True), t)
    mu = K.dot(a)
    Z = la.solve_triangular(L, Ss[:, None] * K, lower=True).T       # Z = K S^1/2 L^-T
    return mu, Z


def ep_approximation(K, y, calc_cov=True, tol=1e-6, max_iters=100, damping=1.0):
    """Parallel EP.  Returns (f_post, C, n_cubic_ops) like post_approx_func (estimators.py:126-139), or (f_post, ops)
    without calc_cov; ops = iterations (+1 with calc_cov).  Raises after max_iters like lpa.py:100-102."""
    n = y.shape[0]
    nu, tau, mu, s2 = np.zeros(n), np.zeros(n), np.zeros(n), np.diag(K).copy()
    it, done = 0, False
    Z = None
    while not done and it < max_iters:
        tau_c = 1. / s2 - tau
        nu_c = mu / s2 - nu
        mu_h, s2_h = _probit_moments(y, nu_c / tau_c, 1. / tau_c)
        tau_n = np.maximum(tau + damping * ((1. / s2_h - tau_c) - tau), 0.)
        nu_n = nu + damping * ((mu_h / s2_h - nu_c) - nu)
        delta = max(np.max(np.abs(tau_n - tau)), np.max(np.abs(nu_n - nu)))
        tau, nu = tau_n, nu_n
        mu, Z = _ep_posterior(K, nu, tau)
        s2 = np.diag(K) - (Z * Z).sum(1)
        it += 1
        done = delta < tol
    if not done:
        raise MaximumIterationsExceededError('Failed to converge in {0} iterations'.format(it))
    if calc_cov:
        return mu, K - Z.dot(Z.T), it + 1
    return mu, it


def ep_sequential_textbook(K, y, tol=1e-8, max_sweeps=100):
    """GPML Alg. 3.5 as printed: sites visited in order with rank-one updates of Sigma, Sigma and mu recomputed from
    the Cholesky factor of B after every sweep.  O(n^3) per sweep in Python loops: small n only."""
    n = y.shape[0]
    nu, tau, mu, Sigma = np.zeros(n), np.zeros(n), np.zeros(n), K.copy()
    for sweep in range(max_sweeps):
        nu_old, tau_old = nu.copy(), tau.copy()
        for i in range(n):
            tau_c = 1. / Sigma[i, i] - tau[i]
            nu_c = mu[i] / Sigma[i, i] - nu[i]
            mu_h, s2_h = _probit_moments(y[i], nu_c / tau_c, 1. / tau_c)
            dt = 1. / s2_h - tau_c - tau[i]
            tau[i] += dt
            nu[i] = mu_h / s2_h - nu_c
            si = Sigma[:, i].copy()
            Sigma -= (dt / (1. + dt * si[i])) * np.outer(si, si)
            mu = Sigma.dot(nu)
        mu, Z = _ep_posterior(K, nu, tau)
        Sigma = K - Z.dot(Z.T)
        if max(np.max(np.abs(tau - tau_old)), np.max(np.abs(nu - nu_old))) < tol:
            return mu, Sigma, sweep + 1
    raise MaximumIterationsExceededError('Failed to converge in {0} sweeps'.format(max_sweeps))


def kernel_gradients(X, theta, ard):
    """dK/dtheta_p, shape (n_theta, n, n).  NO REFERENCE COUNTERPART (the reference has no gradients):
    this is the analytic derivative of the two builders above, checked against central differences of them in
    tests/test_oracle_vs_golden.py; it exists only to check the CUDA extension apm_kernel_grad."""
    X = np.asarray(X, dtype=np.float64)
    theta = np.asarray(theta, dtype=np.float64)
    n, D = X.shape
    sigma = np.exp(theta[0])
    diff = X[:, None, :] - X[None, :, :]
    off = 1. - np.eye(n)
    if ard:
        d2 = (diff / np.exp(theta[1:]))**2
        K = sigma * np.exp(-0.5 * d2.sum(-1))
        out = np.empty((D + 1, n, n))
        out[0] = K
        for k in range(D):
            out[k + 1] = K * d2[:, :, k] * off
    else:
        r2 = (diff**2).sum(-1)
        tau = np.exp(theta[1])
        K = sigma * np.exp(-r2 / (2. * tau**2))
        out = np.empty((2, n, n))
        out[0] = K
        out[1] = K * r2 / tau**2 * off
    return out


# ----------------------------------------------------------------------------------------------
# Laplace approximation  (gpdemo/latent_posterior_approximations.py:22-124)
# ----------------------------------------------------------------------------------------------

def laplace_approximation(K, y, calc_cov=True, calc_lml=False, diff_f_tol=1e-4, max_iters=1000):
    """Newton mode search in the GPML Alg. 3.1 form (lpa.py:81-99) then optional log-ML
    (lpa.py:103-106) and covariance (lpa.py:107-112).  Quirks kept: starts from f = 0; stops on
    mean squared change < tol; covariance / lml use L, W^1/2 K and `a` of the LAST executed
    iteration (evaluated at the previous f); returned op count is iters+1 with calc_cov."""
    n = y.shape[0]
    f = np.zeros(n)
    it = 0
    done = False
    while not done and it < max_iters:
        v = np.exp(-0.5 * f**2 - log_ndtr(y * f) - HALF_LOG_2PI)     # lpa.py:86
        g = v * y                                                    # lpa.py:87
        W = v**2 + g * f                                             # lpa.py:88
        Ws = W**0.5                                                  # lpa.py:89
        WsK = (Ws * K).T                   # [i,j] = Ws[i] K[j,i]      lpa.py:90
        B = np.eye(n) + WsK * Ws                                     # lpa.py:91
        L = la.cholesky(B, lower=True)                               # lpa.py:92
        b = W * f + g                                                # lpa.py:93
        a = b - Ws * la.cho_solve((L, True), WsK.dot(b))             # lpa.py:94
        f_new = K.dot(a)                                             # lpa.py:95
        done = np.mean((f_new - f)**2) < diff_f_tol                  # lpa.py:96-97
        f = f_new
        it += 1
    if not done:
        raise MaximumIterationsExceededError('Failed to converge in {0} iterations'.format(it))
    out = [f]
    if calc_cov:
        C = K - WsK.T.dot(la.cho_solve((L, True), WsK))              # lpa.py:111-112
        out.append(C)
    if calc_lml:
        lml = -0.5 * a.dot(f) + log_ndtr(y * f).sum() - np.log(L.diagonal()).sum()  # lpa.py:105-106
        out.append(lml)
    out.append(it + 1 if calc_cov else it)                           # lpa.py:113-124
    return tuple(out)


# ----------------------------------------------------------------------------------------------
# estimators  (gpdemo/estimators.py)
# ----------------------------------------------------------------------------------------------

class LogMarginalLikelihoodLaplaceEstimator(object):
    """gpdemo/estimators.py:19-82."""

    def __init__(self, X, y, kernel_func):
        self.X, self.y, self.kernel_func = X, y, kernel_func
        self._K = np.empty((X.shape[0], X.shape[0]))
        self.n_cubic_ops = 0

    def reset_cubic_op_count(self):
        self.n_cubic_ops = 0

    def __call__(self, theta):
        self.kernel_func(self._K, self.X, theta)                         # est.py:78
        f, lml, ops = laplace_approximation(self._K, self.y, calc_cov=False, calc_lml=True)
        self.n_cubic_ops += ops                                          # est.py:81
        return lml


def is_log_weights(ns, y, K_chol, C_chol, f_post):
    """Per-importance-sample log weights, gpdemo/estimators.py:221-238 (before the logsumexp)."""
    f_s = f_post[None] + C_chol.dot(ns).T                                         # est.py:223
    qK = (la.cho_solve((K_chol, True), f_s.T) * f_s.T).sum(0)                     # est.py:225
    log_prior_f = -0.5 * qK - np.log(K_chol.diagonal()).sum()                     # est.py:226-227
    log_lik = log_ndtr(f_s * y).sum(-1)                                           # est.py:229
    zm = f_s - f_post[None]                                                       # est.py:232
    qC = (la.cho_solve((C_chol, True), zm.T) * zm.T).sum(0)                       # est.py:233-234
    log_q = -0.5 * qC - np.log(C_chol.diagonal()).sum()                           # est.py:235-236
    return log_lik + log_prior_f - log_q                                          # est.py:238-239


class LogMarginalLikelihoodApproxPosteriorISEstimator(object):
    """gpdemo/estimators.py:90-241."""

    def __init__(self, X, y, kernel_func, post_approx_func):
        self.X, self.y = X, y
        self.kernel_func, self.post_approx_func = kernel_func, post_approx_func
        self._K = np.empty((X.shape[0], X.shape[0]))
        self.n_cubic_ops = 0

    def reset_cubic_op_count(self):
        self.n_cubic_ops = 0

    def __call__(self, ns, theta=None, cached_results=None):
        if theta is None and cached_results is None:                      # est.py:201-202
            raise ValueError('One of theta or cached_results must be provided')
        if cached_results is None:                                        # est.py:203-217
            self.kernel_func(self._K, self.X, theta)
            K_chol = la.cholesky(self._K, lower=True)
            f_post, C, ops = self.post_approx_func(self._K, self.y)
            try:
                C_chol = la.cholesky(C, lower=True)
            except la.LinAlgError:
                e = la.eigvalsh(C)
                raise InvalidCovarianceMatrixError(
                    'Posterior covariance matrix not PSD: sum of negative eigenvalues {0}'
                    .format(e[e <= 0].sum()))
            self.n_cubic_ops += ops + 2
        else:
            K_chol, C_chol, f_post = cached_results                       # est.py:218-220
        lw = is_log_weights(ns, self.y, K_chol, C_chol, f_post)
        return logsumexp(lw) - np.log(ns.shape[1]), (K_chol, C_chol, f_post)   # est.py:240-241


class LogMarginalLikelihoodPriorMCEstimator(object):
    """gpdemo/estimators.py:244-325."""

    def __init__(self, X, y, kernel_func):
        self.X, self.y, self.kernel_func = X, y, kernel_func
        self._K = np.empty((X.shape[0], X.shape[0]))
        self.n_cubic_ops = 0

    def reset_cubic_op_count(self):
        self.n_cubic_ops = 0

    def __call__(self, ns, theta=None, K_chol=None):
        if theta is None and K_chol is None:                              # est.py:317-318
            raise ValueError('One of theta or K_chol must be provided')
        if K_chol is None:
            self.kernel_func(self._K, self.X, theta)
            K_chol = la.cholesky(self._K, lower=True)
            self.n_cubic_ops += 1                                         # est.py:322
        f_s = K_chol.dot(ns).T                                            # est.py:323
        ll = log_ndtr(f_s * self.y[None]).sum(-1)                         # est.py:324
        return logsumexp(ll) - np.log(ns.shape[1]), K_chol               # est.py:325


# ----------------------------------------------------------------------------------------------
# helpers  (gpdemo/utils.py)
# ----------------------------------------------------------------------------------------------

def log_gamma_log_pdf(x, a, b):
    """gpdemo/utils.py:39-59: density of x when exp(x) ~ Gamma(shape a, rate b)."""
    return a * np.log(b) - gammaln(a) + a * x - b * np.exp(x)


def adapt_factor_func(b, n_batch):
    """gpdemo/utils.py:62-83."""
    return 5. - min(b + 1, n_batch / 5.) / (n_batch / 5.) * 3.9


def normalise_inputs(X):
    """gpdemo/utils.py:86-105."""
    mn, sd = X.mean(0), X.std(0)
    return (X - mn[None]) / sd[None], mn, sd
