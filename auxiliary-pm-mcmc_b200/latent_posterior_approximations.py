"""Drop-in for `gpdemo.latent_posterior_approximations` (lpa.py): `laplace_approximation(K, y, ...)`
with the reference's signature, return arity (lpa.py:113-124) and exception, computed on the GPU
(Newton iterations with the blocked DMMA Cholesky of B = I + W^1/2 K W^1/2, lpa.py:85-99; covariance
lpa.py:107-112)."""
import numpy as np

from . import _capi


class MaximumIterationsExceededError(Exception):
    """Newton's method did not converge within max_iters (lpa.py:17-19)."""


_engines = {}


def _engine_for(y):
    y = np.asarray(y, dtype=np.float64)
    key = (y.shape[0], y.tobytes())
    eng = _engines.get(key)
    if eng is None:
        if len(_engines) >= 4:
            _engines.pop(next(iter(_engines))).close()
        eng = _capi.Engine(np.zeros((y.shape[0], 1)), y, kernel='iso', max_chains=1, n_slots=1, max_nimp=1)
        _engines[key] = eng
    return eng


def raise_for_status(status, iters=None):
    """Turn a per-chain status code of the C ABI into the reference's exception."""
    if status == _capi.CHAIN_OK:
        return
    if status == _capi.CHAIN_NEWTON_MAXIT:
        raise MaximumIterationsExceededError('Failed to converge in {0} iterations'.format(iters))
    if status in (_capi.CHAIN_CHOL_K, _capi.CHAIN_CHOL_B):
        raise np.linalg.LinAlgError('matrix is not positive definite (Cholesky failed on the device)')
    if status == _capi.CHAIN_CHOL_C:
        from .estimators import InvalidCovarianceMatrixError
        raise InvalidCovarianceMatrixError('Posterior covariance matrix not PSD')
    raise ValueError('array must not contain infs or NaNs')


def laplace_approximation(K, y, calc_cov=True, calc_lml=False, diff_f_tol=1e-4, max_iters=1000):
    """Gaussian (Laplace) approximation to p(f | y, theta) for the probit likelihood.

    Returns, as the reference does: (f, C, lml, ops) / (f, lml, ops) / (f, C, ops) / (f, ops) depending
    on the two flags, with ops = Newton iterations (+1 with calc_cov)."""
    K = np.asarray(K, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if not np.isfinite(K).all():
        raise ValueError('array must not contain infs or NaNs')
    eng = _engine_for(y)
    eng.set_newton(diff_f_tol, max_iters)
    f, C, lml, ops, st = eng.laplace(K, calc_cov=calc_cov, calc_lml=calc_lml)
    raise_for_status(int(st[0]), int(ops[0]))
    out = [f[0]]
    if calc_cov:
        out.append(C[0])
    if calc_lml:
        out.append(float(lml[0]))
    out.append(int(ops[0]))
    return tuple(out)


def ep_approximation(K, y, calc_cov=True, tol=1e-6, max_iters=100, damping=1.0):
    """EXTENSION (the reference has only laplace_approximation): expectation-propagation approximation to
    p(f | y, theta) for the probit likelihood, usable wherever the reference takes a
    `post_approx_func(K, y) -> (f_post, C, cubic_ops)` (estimators.py:126-139).  GPML Alg. 3.5 with all sites updated
    per sweep, on the GPU (run_ep in csrc/apm_capi.cu; CPU restatement: the ep_approximation restatement under oracle/).

    Returns (f, C, ops) or (f, ops) without calc_cov; ops = EP iterations (+1 with calc_cov).  Raises
    MaximumIterationsExceededError like the Laplace approximation does."""
    K = np.asarray(K, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if not np.isfinite(K).all():
        raise ValueError('array must not contain infs or NaNs')
    eng = _engine_for(y)
    eng.set_approximation('ep', tol, max_iters, damping)
    try:
        f, C, nu, tau, ops, st = eng.ep(K, calc_cov=calc_cov)
    finally:
        eng.set_approximation('laplace')
    raise_for_status(int(st[0]), int(ops[0]))
    if calc_cov:
        return f[0], C[0], int(ops[0])
    return f[0], int(ops[0])
