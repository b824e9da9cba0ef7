"""apm_b200.samplers / mcmc_updates (host-side mirrors of auxpm) driven by the CPU oracle estimator must
reproduce the reference's chains: same theta trace, reject counts and cubic-op counts for a fixed seed,
over all 1000 iterations of the golden runs (oracle/gen_golden.py ran the reference's own samplers)."""
import warnings

import numpy as np
import pytest

import apm_oracle as orc
from apm_b200 import samplers as smp, utils
from conftest import load_golden
from wiring import run_golden_case, build_sampler, first_divergence

ORACLE_IMPL = dict(est_cls=orc.LogMarginalLikelihoodApproxPosteriorISEstimator, lap_func=orc.laplace_approximation,
                   iso_kernel=orc.isotropic_squared_exponential_kernel, log_gamma_log_pdf=utils.log_gamma_log_pdf,
                   smp=smp)


@pytest.mark.parametrize('method', ['mi+mh', 'ess+mh', 'mi+rdss', 'ess+rdss', 'pmmh'])
@pytest.mark.parametrize('N', [1, 4])
def test_chain_matches_reference(method, N):
    g = load_golden('samplers')
    n_iter = int(g['n_iter'])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        thetas, n_rej, ops = run_golden_case(method, N, g, n_iter, **ORACLE_IMPL)
    key = '%s_N%d_' % (method, N)
    assert first_divergence(thetas, g[key + 'thetas']) is None
    assert np.array_equal(n_rej, g[key + 'n_reject'])
    assert ops == int(g[key + 'cubic_ops'])


def test_adaptive_run_matches_reference():
    g = load_golden('samplers')
    prng = np.random.RandomState()
    s, ml = build_sampler('mi+mh', g['X'], g['y'], 1, prng, **ORACLE_IMPL)
    prng.seed(4242)
    from apm_b200 import synth
    theta_init = synth.draw_theta_prior(prng, g['X'].shape[1], ard=False)
    s.prop_scales = np.array([0.5, 0.5])
    th, sc, acc = s.adaptive_run(theta_init, 25, 8, 0.15, 0.30, utils.adapt_factor_func)
    assert first_divergence(th, g['adapt_thetas']) is None
    np.testing.assert_allclose(sc, g['adapt_scales'], rtol=1e-12)
    np.testing.assert_allclose(acc, g['adapt_accept'], rtol=1e-12)


def test_utils_match_reference():
    g = load_golden('utils')
    assert np.array_equal(utils.log_gamma_log_pdf(g['xs'], 1.1, 0.1), g['lg_11_01'])
    assert np.array_equal(np.array([utils.adapt_factor_func(b, 20) for b in range(20)]), g['adapt'])
    X = np.random.RandomState(0).normal(size=(20, 3)) * 3 + 1
    Xn, mn, sd = utils.normalise_inputs(X)
    np.testing.assert_allclose(Xn.mean(0), 0, atol=1e-14)
    np.testing.assert_allclose(Xn.std(0), 1, atol=1e-14)
