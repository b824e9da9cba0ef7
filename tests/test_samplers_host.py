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


def test_headline_shape_chain_matches_reference():
    """The oracle restatement under the host samplers reproduces the reference's own chain at the headline shape
    (pima-shaped n = 768, D = 8, N_imp = 64, E-SS u + RD-SS theta): first 12 iterations of the golden trace."""
    g = load_golden('samplers_fullsize')
    n_iter = 12
    gg = dict(X=g['pima_X'], y=g['pima_y'])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        thetas, n_rej, ops = run_golden_case('ess+rdss', 64, gg, n_iter, **ORACLE_IMPL)
    assert first_divergence(thetas, g['pima_ess+rdss_N64_thetas'][:n_iter]) is None


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


def test_run_artefacts_have_the_reference_file_format(tmp_path):
    """save_run / save_adaptive_run (gpdemo/utils.py:108-208): key names, perf-stat packing and sorted JSON."""
    import json
    thetas = np.arange(12.).reshape(6, 2)
    res, par = utils.save_run(str(tmp_path), 'apm_test', thetas, (3, 4), 77, 1.5, {'n_imp': 4, 'a': [1, 2]})
    z = np.load(res)
    assert sorted(z.files) == ['n_reject_n_cubic_ops_comp_time', 'thetas']
    assert np.array_equal(z['thetas'], thetas) and np.array_equal(z['n_reject_n_cubic_ops_comp_time'], [3, 4, 77, 1.5])
    assert res.endswith('apm_test_results.npz') and par.endswith('apm_test_params.json')
    assert json.load(open(par)) == {'n_imp': 4, 'a': [1, 2]} and open(par).read().index('"a"') < open(par).read().index('"n_imp"')
    res, par = utils.save_adaptive_run(str(tmp_path), 'ad', thetas, thetas[:3], np.ones(3), thetas, 5, 9, 2.5, {})
    z = np.load(res)
    assert sorted(z.files) == ['adapt_accept_rates', 'adapt_prop_scales', 'adapt_thetas', 'n_reject_n_cubic_ops_comp_time', 'thetas']
    assert np.array_equal(z['n_reject_n_cubic_ops_comp_time'], [5, 9, 2.5])


def test_run_artefacts_match_files_written_by_the_reference(tmp_path):
    """tests/golden/run_artefacts/ref_* were written by the unmodified gpdemo.utils.save_run / save_adaptive_run
    (oracle/gen_golden.py:gen_run_artefacts): same npz members (names, dtypes, shapes, values) and byte-identical JSON."""
    import os
    gdir = os.path.join(os.path.dirname(__file__), 'golden', 'run_artefacts')
    thetas = np.arange(12.).reshape(6, 2)
    made = {
        'apm_test': utils.save_run(str(tmp_path), 'apm_test', thetas, (3, 4), 77, 1.5, {'n_imp': 4, 'a': [1, 2], 'tag': 'x'}),
        'pmmh_test': utils.save_run(str(tmp_path), 'pmmh_test', thetas, 5, 9, 2.5, {'seed': 1}),
        'ad_test': utils.save_adaptive_run(str(tmp_path), 'ad_test', thetas, thetas[:3], np.ones(3) * 0.25, thetas + 1, (1, 2), 9,
                                           2.5, {'n_batch': 3, 'batch_size': 2}),
    }
    for tag, (res, par) in made.items():
        ref = np.load(os.path.join(gdir, 'ref_%s_results.npz' % tag))
        ours = np.load(res)
        assert sorted(ours.files) == sorted(ref.files)
        for k in ref.files:
            assert ours[k].dtype == ref[k].dtype and ours[k].shape == ref[k].shape and np.array_equal(ours[k], ref[k]), (tag, k)
        assert open(par).read() == open(os.path.join(gdir, 'ref_%s_params.json' % tag)).read()
        # <time stamp><tag>_results.npz / _params.json with the reference's stamp format
        base = os.path.basename(res)
        assert base.endswith(tag + '_results.npz') and len(base) == len('YYYY_mm_dd_HH_MM_SS_') + len(tag + '_results.npz')


def test_chain_diagnostics():
    rs = np.random.RandomState(3)
    n = 20000
    white = rs.normal(size=n)
    assert abs(utils.effective_sample_size(white) / n - 1.) < 0.1
    rho = 0.9                                              # AR(1): ESS = n (1 - rho) / (1 + rho)
    ar = np.empty(n)
    ar[0] = white[0]
    for t in range(1, n):
        ar[t] = rho * ar[t - 1] + white[t]
    assert abs(utils.effective_sample_size(ar) / (n * (1 - rho) / (1 + rho)) - 1.) < 0.2
    same = rs.normal(size=(4, 2000))
    assert abs(utils.gelman_rubin(same) - 1.) < 0.01
    assert utils.gelman_rubin(same + np.arange(4)[:, None] * 3.) > 2.
