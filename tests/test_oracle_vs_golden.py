"""The CPU oracle (oracle/apm_oracle.py) against golden vectors produced by the UNMODIFIED reference
(oracle/gen_golden.py).  This is what pins the oracle: the reference has no tests of its own."""
import numpy as np
import pytest

import apm_oracle as orc
from conftest import load_golden

RTOL_LA = 1e-11   # quantities that go through LAPACK (thread count / kernel choice can move last bits)


def test_kernels_bit_exact():
    g = load_golden('kernels')
    for tag in 'ab':
        X = g['X_' + tag]
        n = X.shape[0]
        for t in range(3):
            K = np.empty((n, n))
            orc.isotropic_squared_exponential_kernel(K, X, g['th_iso_' + tag][t], float(g['eps_iso']))
            assert np.array_equal(K, g['K_iso_' + tag][t])
            orc.diagonal_squared_exponential_kernel(K, X, g['th_ard_' + tag][t], float(g['eps_ard']))
            assert np.array_equal(K, g['K_ard_' + tag][t])


def test_laplace():
    g = load_golden('laplace')
    for tag in 'ab':
        K, y = g['K_' + tag], g['y_' + tag]
        f, C, lml, ops = orc.laplace_approximation(K, y, calc_cov=True, calc_lml=True)
        np.testing.assert_allclose(f, g['f_' + tag], rtol=RTOL_LA, atol=1e-13)
        np.testing.assert_allclose(C, g['C_' + tag], rtol=RTOL_LA, atol=1e-13)
        assert abs(lml - g['lml_' + tag]) <= RTOL_LA * abs(g['lml_' + tag])
        assert ops == int(g['ops_cov_' + tag])
        f2, lml2, ops2 = orc.laplace_approximation(K, y, calc_cov=False, calc_lml=True)
        assert ops2 == int(g['ops_nocov_' + tag]) == ops - 1
        f3, ops3 = orc.laplace_approximation(K, y, calc_cov=False)
        assert ops3 == ops2 and np.array_equal(f2, f3)


def test_laplace_max_iters():
    g = load_golden('laplace')
    with pytest.raises(orc.MaximumIterationsExceededError):
        orc.laplace_approximation(g['K_a'], g['y_a'], max_iters=1)


@pytest.mark.parametrize('name', ['small_ard', 'small_iso', 'pima_ard', 'pima_iso', 'breast_ard'])
def test_estimators(name):
    g = load_golden('estimator_' + name)
    X, y, thetas, kind = g['X'], g['y'], g['thetas'], str(g['kind'])
    n = X.shape[0]
    base = orc.diagonal_squared_exponential_kernel if kind == 'ard' else orc.isotropic_squared_exponential_kernel
    kf = lambda K, X_, th: base(K, X_, th, float(g['eps']))  # noqa: E731
    n_theta = thetas.shape[0] if n < 500 else 1     # keep the CPU suite short at the pima/breast shapes
    for t in range(n_theta):
        for N in g['Ns']:
            est = orc.LogMarginalLikelihoodApproxPosteriorISEstimator(X, y, kf, orc.laplace_approximation)
            u1 = np.random.RandomState(7000 + 10 * t + int(N)).normal(size=(n, int(N)))
            u2 = np.random.RandomState(8000 + 10 * t + int(N)).normal(size=(n, int(N)))
            full, cache = est(u1, thetas[t])
            cached, _ = est(u2, None, cache)
            key = 't%d_N%d_' % (t, N)
            assert abs(full - g[key + 'full']) <= 1e-10 * abs(g[key + 'full'])
            assert abs(cached - g[key + 'cached']) <= 1e-10 * abs(g[key + 'cached'])
            assert est.n_cubic_ops == int(g[key + 'cubic_ops'])
        np.testing.assert_allclose(cache[2], g['t%d_f_post' % t], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(cache[0].diagonal(), g['t%d_diagK' % t], rtol=1e-9)
        np.testing.assert_allclose(cache[1].diagonal(), g['t%d_diagC' % t], rtol=1e-9)
        lap = orc.LogMarginalLikelihoodLaplaceEstimator(X, y, kf)
        assert abs(lap(thetas[t]) - g['t%d_laplace_lml' % t]) <= 1e-10 * abs(g['t%d_laplace_lml' % t])
        assert lap.n_cubic_ops == int(g['t%d_laplace_ops' % t])
        pm = orc.LogMarginalLikelihoodPriorMCEstimator(X, y, kf)
        u3 = np.random.RandomState(9000 + t).normal(size=(n, int(g['Ns'][-1])))
        v, _ = pm(u3, thetas[t])
        assert abs(v - g['t%d_prior_mc' % t]) <= 1e-10 * abs(g['t%d_prior_mc' % t])


def test_estimator_errors():
    g = load_golden('estimator_small_ard')
    est = orc.LogMarginalLikelihoodApproxPosteriorISEstimator(
        g['X'], g['y'], orc.diagonal_squared_exponential_kernel, orc.laplace_approximation)
    with pytest.raises(ValueError):
        est(np.zeros((g['X'].shape[0], 1)))


def test_utils():
    g = load_golden('utils')
    assert np.array_equal(orc.log_gamma_log_pdf(g['xs'], 1.1, 0.1), g['lg_11_01'])
    assert np.array_equal(orc.log_gamma_log_pdf(g['xs'], 1., 0.3), g['lg_1_03'])
    assert np.array_equal(np.array([orc.adapt_factor_func(b, 20) for b in range(20)]), g['adapt'])


def test_oracle_kernel_gradients_vs_central_differences():
    """The gradient restatement has no reference counterpart (the reference has no gradients): it is pinned to
    central differences of the golden-pinned kernel builders instead."""
    rs = np.random.RandomState(0)
    n, D = 23, 3
    X = rs.normal(size=(n, D))
    for ard, theta in ((True, np.array([0.3, 0.2, -0.1, 0.5])), (False, np.array([-0.2, 0.4]))):
        build = orc.diagonal_squared_exponential_kernel if ard else orc.isotropic_squared_exponential_kernel
        g = orc.kernel_gradients(X, theta, ard)
        h = 1e-5
        for p in range(theta.shape[0]):
            Kp, Km = np.empty((n, n)), np.empty((n, n))
            tp, tm = theta.copy(), theta.copy()
            tp[p] += h
            tm[p] -= h
            build(Kp, X, tp, 1e-8)
            build(Km, X, tm, 1e-8)
            fd = (Kp - Km) / (2 * h)
            assert np.max(np.abs(fd - g[p])) < 1e-8 * max(1., np.max(np.abs(g[p])))


def test_oracle_parallel_ep_matches_textbook_sequential_ep():
    """EP has no reference counterpart: the parallel-EP restatement (what the CUDA path runs) is pinned to the
    textbook sequential sweep of GPML Alg. 3.5 -- same fixed point."""
    from apm_b200 import synth
    for n, D, shift in ((60, 2, 0.), (90, 3, 0.4)):
        X, y, th = synth.make_dataset(n, D, seed=4)
        K = np.empty((n, n))
        orc.diagonal_squared_exponential_kernel(K, X, th + shift, 1e-8)
        mu_p, C_p, ops = orc.ep_approximation(K, y, tol=1e-10)
        mu_s, C_s, sweeps = orc.ep_sequential_textbook(K, y, tol=1e-10)
        assert np.max(np.abs(mu_p - mu_s)) < 1e-8 and np.max(np.abs(C_p - C_s)) < 1e-8
        assert 3 <= ops <= 40
        # EP matches the first two moments better than Laplace: its mean differs from the Laplace mode
        f_l = orc.laplace_approximation(K, y, calc_cov=False)[0]
        assert np.max(np.abs(f_l - mu_p)) > 1e-3


def test_estimator_newton_iteration_counts():
    """Oracle vs the reference at the headline shape for thetas on which the reference's Newton loop takes 3 / 5 / 6
    iterations (estimator_pima_iters.npz): the restatement stops at the same iteration (Quirk 3)."""
    g = load_golden('estimator_pima_iters')
    X, y, thetas, N = g['X'], g['y'], g['thetas'], int(g['N'])
    n = X.shape[0]
    kf = lambda K, X_, th: orc.diagonal_squared_exponential_kernel(K, X_, th, float(g['eps']))  # noqa: E731
    iters = [int(v) for v in g['newton_iters']]
    picks = [iters.index(3), iters.index(5), iters.index(6)]       # one theta per count keeps the CPU suite short
    for t in picks:
        est = orc.LogMarginalLikelihoodApproxPosteriorISEstimator(X, y, kf, orc.laplace_approximation)
        u1 = np.random.RandomState(7100 + t).normal(size=(n, N))
        full, _ = est(u1, thetas[t])
        key = 't%d_' % t
        assert est.n_cubic_ops == int(g[key + 'cubic_ops']) == iters[t] + 3
        assert abs(full - g[key + 'full']) <= max(1e-10 * abs(g[key + 'full']), 10. * float(g[key + 'ulp_sens']))


@pytest.mark.parametrize('method', ['ess+rdss', 'mi+mh', 'ess+mh', 'mi+rdss', 'pmmh'])
def test_sampler_restatement_vs_reference_chains(method):
    """oracle/apm_oracle_samplers.py (the chain loops timed by bench.py's reference arm) against the 1000-iteration chains
    of the UNMODIFIED reference (tests/golden/samplers.npz): identical traces, reject counts and cubic-op counts."""
    import apm_oracle_samplers as osm
    from apm_b200 import synth
    from wiring import first_divergence
    g = load_golden('samplers')
    X, y = g['X'], g['y']
    N, n_iter = 4, 400
    D = X.shape[1]
    prior = synth.prior_params(D)
    kf = lambda K, X_, th: orc.isotropic_squared_exponential_kernel(K, X_, th, 1e-8)  # noqa: E731
    ml = orc.LogMarginalLikelihoodApproxPosteriorISEstimator(X, y, kf, orc.laplace_approximation)
    lg = orc.log_gamma_log_pdf
    log_prior = lambda th: lg(th[0], prior['a_sigma'], prior['b_sigma']) + lg(th[1], prior['a_tau'], prior['b_tau'])  # noqa: E731

    def log_f_estimator(u, theta=None, cached=None):
        v, c = ml(u, theta, cached)
        return v + log_prior(theta), c

    prng = np.random.RandomState()
    u_sampler = lambda: prng.normal(size=(y.shape[0], N))  # noqa: E731
    prop_sampler = lambda th, s: np.r_[th[0] + s[0] * prng.normal(), th[1] + s[1] * prng.normal()]  # noqa: E731
    log_prop_density = lambda tp, tc, s: -0.5 * (((tp[0] - tc[0]) / s[0])**2 + ((tp[1] - tc[1]) / s[1])**2)  # noqa: E731

    def dir_and_w():
        d = prng.normal(size=2)
        d /= d.dot(d)**0.5
        return d, 1.

    prng.seed(1000 + N)
    theta_init = synth.draw_theta_prior(prng, D, ard=False)
    scales = np.array([0.5, 0.5])
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        if method == 'pmmh':
            main = lambda th: ml(prng.normal(size=(y.shape[0], N)), th)[0] + log_prior(th)  # noqa: E731
            thetas, n_rej = osm.pmmh_chain(main, log_prop_density, prop_sampler, scales, prng, theta_init, n_iter)
        else:
            thetas, n_rej = osm.apm_chain(method, log_f_estimator, u_sampler, prng, theta_init, n_iter, dir_and_w_sampler=dir_and_w,
                                          log_prop_density=log_prop_density, prop_sampler=prop_sampler, prop_scales=scales)
    key = '%s_N%d_' % (method, N)
    assert first_divergence(thetas, g[key + 'thetas'][:n_iter]) is None
