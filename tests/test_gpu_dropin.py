"""The reference-facing Python layer on the GPU: drop-in callables with the reference's signatures,
wired exactly as the notebooks wire them, compared with golden chains of the unmodified reference."""
import functools
import warnings

import numpy as np
import pytest

import apm_oracle as orc
from apm_b200 import estimators as est, kernels as krn, latent_posterior_approximations as lpa
from apm_b200 import samplers as smp, utils, synth
from conftest import load_golden
from wiring import run_golden_case, first_divergence

pytestmark = pytest.mark.gpu

GPU_IMPL = dict(est_cls=est.LogMarginalLikelihoodApproxPosteriorISEstimator, lap_func=lpa.laplace_approximation,
                iso_kernel=krn.isotropic_squared_exponential_kernel, log_gamma_log_pdf=utils.log_gamma_log_pdf, smp=smp)


def test_kernel_functions_in_place():
    g = load_golden('kernels')
    X = g['X_b']
    n = X.shape[0]
    K = np.empty((n, n))
    assert krn.isotropic_squared_exponential_kernel(K, X, g['th_iso_b'][1], float(g['eps_iso'])) is None
    assert np.max(np.abs(K - g['K_iso_b'][1]) / g['K_iso_b'][1]) < 4.5e-16
    Kf = np.empty((n, n), order='F')
    krn.diagonal_squared_exponential_kernel(Kf, X, g['th_ard_b'][2], epsilon=float(g['eps_ard']))
    assert np.max(np.abs(Kf - g['K_ard_b'][2]) / g['K_ard_b'][2]) < 4.5e-16
    with pytest.raises(ValueError):
        krn.isotropic_squared_exponential_kernel(K, X, np.zeros(3))


def test_laplace_function_return_arity():
    g = load_golden('laplace')
    K, y = g['K_b'], g['y_b']
    f, C, lml, ops = lpa.laplace_approximation(K, y, calc_cov=True, calc_lml=True)
    assert ops == int(g['ops_cov_b']) and abs(lml - g['lml_b']) < 1e-10 * abs(g['lml_b'])
    np.testing.assert_allclose(C, g['C_b'], rtol=1e-9, atol=1e-12)
    f2, lml2, ops2 = lpa.laplace_approximation(K, y, calc_cov=False, calc_lml=True)
    f3, C3, ops3 = lpa.laplace_approximation(K, y)
    f4, ops4 = lpa.laplace_approximation(K, y, calc_cov=False)
    assert (ops2, ops3, ops4) == (ops - 1, ops, ops - 1)
    np.testing.assert_allclose(f, g['f_b'], rtol=1e-10, atol=1e-13)
    with pytest.raises(lpa.MaximumIterationsExceededError):
        lpa.laplace_approximation(K, y, max_iters=1)
    with pytest.raises(ValueError):
        lpa.laplace_approximation(K * np.nan, y)


@pytest.mark.parametrize('name', ['small_iso', 'pima_ard'])
def test_estimator_classes_vs_golden(name):
    g = load_golden('estimator_' + name)
    X, y, thetas, kind = g['X'], g['y'], g['thetas'], str(g['kind'])
    n = X.shape[0]
    base = krn.diagonal_squared_exponential_kernel if kind == 'ard' else krn.isotropic_squared_exponential_kernel
    kf = lambda K, X_, th: base(K, X_, th, float(g['eps']))  # noqa: E731  (the notebooks' wrapper)
    for t in range(2):
        for N in [int(v) for v in g['Ns']]:
            e = est.LogMarginalLikelihoodApproxPosteriorISEstimator(X, y, kf, lpa.laplace_approximation)
            u1 = np.random.RandomState(7000 + 10 * t + N).normal(size=(n, N))
            u2 = np.random.RandomState(8000 + 10 * t + N).normal(size=(n, N))
            full, cache = e(u1, thetas[t])
            cached, cache2 = e(u2, None, cache)
            key = 't%d_N%d_' % (t, N)
            assert isinstance(full, float) and cache2 is cache
            assert abs(full - g[key + 'full']) < 1e-10 * abs(g[key + 'full'])
            assert abs(cached - g[key + 'cached']) < 1e-10 * abs(g[key + 'cached'])
            assert e.n_cubic_ops == int(g[key + 'cubic_ops'])
        K_chol, C_chol, f_post = cache                      # unpacks like the reference's tuple
        np.testing.assert_allclose(f_post, g['t%d_f_post' % t], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(K_chol.diagonal(), g['t%d_diagK' % t], rtol=1e-9)
        # a cache made of plain arrays (e.g. produced by the reference) is accepted too
        again, _ = e(u2, None, (K_chol, C_chol, f_post))
        assert abs(again - cached) < 1e-12 * abs(cached)
        lap = est.LogMarginalLikelihoodLaplaceEstimator(X, y, kf)
        assert abs(lap(thetas[t]) - g['t%d_laplace_lml' % t]) < 1e-10 * abs(g['t%d_laplace_lml' % t])
        assert lap.n_cubic_ops == int(g['t%d_laplace_ops' % t])
        pm = est.LogMarginalLikelihoodPriorMCEstimator(X, y, kf)
        u3 = np.random.RandomState(9000 + t).normal(size=(n, int(g['Ns'][-1])))
        v, kc = pm(u3, thetas[t])
        assert abs(v - g['t%d_prior_mc' % t]) < 1e-10 * abs(g['t%d_prior_mc' % t]) and pm.n_cubic_ops == 1
        v2, _ = pm(u3, None, kc)
        assert v2 == v and pm.n_cubic_ops == 1


def test_plugin_callables():
    """Foreign kernel_func / post_approx_func (here: the CPU oracle's) are honoured (estimators.py:47-53,
    126-139); functools.partial of our Laplace function keeps the fused path with its tolerances."""
    g = load_golden('estimator_small_ard')
    X, y, th = g['X'], g['y'], g['thetas'][0]
    n = X.shape[0]
    u = np.random.RandomState(3).normal(size=(n, 5))
    ref_est = orc.LogMarginalLikelihoodApproxPosteriorISEstimator(
        X, y, orc.diagonal_squared_exponential_kernel, orc.laplace_approximation)
    ref, _ = ref_est(u, th)
    for kf, pf in [(orc.diagonal_squared_exponential_kernel, orc.laplace_approximation),
                   (krn.diagonal_squared_exponential_kernel, orc.laplace_approximation),
                   (orc.diagonal_squared_exponential_kernel, lpa.laplace_approximation)]:
        e = est.LogMarginalLikelihoodApproxPosteriorISEstimator(X, y, kf, pf)
        v, cache = e(u, th)
        assert abs(v - ref) < 1e-10 * abs(ref) and e.n_cubic_ops == ref_est.n_cubic_ops
        v2, _ = e(u, None, cache)
        assert abs(v2 - ref) < 1e-10 * abs(ref)
    tight = functools.partial(lpa.laplace_approximation, diff_f_tol=1e-10)
    e = est.LogMarginalLikelihoodApproxPosteriorISEstimator(X, y, krn.diagonal_squared_exponential_kernel, tight)
    r = orc.LogMarginalLikelihoodApproxPosteriorISEstimator(
        X, y, orc.diagonal_squared_exponential_kernel, functools.partial(orc.laplace_approximation, diff_f_tol=1e-10))
    assert abs(e(u, th)[0] - r(u, th)[0]) < 1e-10 * abs(ref) and e.n_cubic_ops == r.n_cubic_ops > ref_est.n_cubic_ops


def test_exceptions_match_reference_types():
    rs = np.random.RandomState(2)
    X = rs.normal(size=(30, 2))
    X[5] = X[2]
    y = np.where(rs.uniform(size=30) < 0.5, 1., -1.)
    kf = lambda K, X_, th: krn.isotropic_squared_exponential_kernel(K, X_, th, 0.)  # noqa: E731
    e = est.LogMarginalLikelihoodApproxPosteriorISEstimator(X, y, kf, lpa.laplace_approximation)
    with pytest.raises(np.linalg.LinAlgError):
        e(rs.normal(size=(30, 1)), np.zeros(2))
    one_iter = functools.partial(lpa.laplace_approximation, max_iters=1)
    e = est.LogMarginalLikelihoodApproxPosteriorISEstimator(
        rs.normal(size=(30, 2)), y, krn.isotropic_squared_exponential_kernel, one_iter)
    with pytest.raises(lpa.MaximumIterationsExceededError):
        e(rs.normal(size=(30, 1)), np.zeros(2))


@pytest.mark.parametrize('method', ['mi+mh', 'ess+mh', 'mi+rdss', 'ess+rdss', 'pmmh'])
@pytest.mark.parametrize('N', [1, 4])
def test_accept_reject_sequences_vs_reference(method, N):
    """North-star: accept/reject sequences for fixed random streams agree with the reference over the
    first 1000 iterations.  Golden chains come from the reference's own samplers + estimator."""
    g = load_golden('samplers')
    n_iter = int(g['n_iter'])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        thetas, n_rej, ops = run_golden_case(method, N, g, n_iter, **GPU_IMPL)
    key = '%s_N%d_' % (method, N)
    div = first_divergence(thetas, g[key + 'thetas'])
    assert div is None, 'chains diverge at iteration %d of %d' % (div, n_iter)
    assert np.array_equal(n_rej, g[key + 'n_reject'])
    assert ops == int(g[key + 'cubic_ops'])
