"""Lock-step batched sampler driver on the CPU (oracle backend): every chain of a batch must reproduce the
single-chain trace for its seed -- the reference's golden chains for the golden seed, and apm_b200.samplers
for the other seeds -- whatever else is in the batch."""
import warnings

import numpy as np
import pytest

import apm_oracle as orc
from apm_b200 import batched, samplers as smp, synth, utils
from conftest import load_golden
from oracle_backend import OracleBackend
from wiring import build_sampler, first_divergence

ORACLE_IMPL = dict(est_cls=orc.LogMarginalLikelihoodApproxPosteriorISEstimator, lap_func=orc.laplace_approximation,
                   iso_kernel=orc.isotropic_squared_exponential_kernel, log_gamma_log_pdf=utils.log_gamma_log_pdf,
                   smp=smp)


def single_chain(method, X, y, N, seed, n_iter):
    prng = np.random.RandomState()
    s, ml = build_sampler(method, X, y, N, prng, **ORACLE_IMPL)
    prng.seed(seed)
    theta_init = synth.draw_theta_prior(prng, X.shape[1], ard=False)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        res = s.get_samples(theta_init, n_iter)
    return (res[0] if isinstance(res, tuple) else res), ml.n_cubic_ops


@pytest.mark.parametrize('method', ['mi+mh', 'ess+mh', 'mi+rdss', 'ess+rdss', 'pmmh'])
def test_batched_chains_match_single_chain_and_reference(method):
    g = load_golden('samplers')
    X, y = g['X'], g['y']
    N, n_iter = 4, 120
    seeds = [1000 + N, 77, 4242]                      # first seed = the golden (reference) chain
    drv = batched.BatchedAPMSampler(OracleBackend(X, y), X.shape[0], N, 2, method, batched.make_log_prior(X.shape[1], False),
                                    seeds, prop_scales=[0.5, 0.5], slice_width=1.)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        out = drv.get_samples(None, n_iter, theta_init_sampler=lambda prng: synth.draw_theta_prior(prng, X.shape[1], ard=False))
    assert np.all(out['failed'] == 0)
    key = '%s_N%d_' % (method, N)
    assert first_divergence(out['thetas'][0], g[key + 'thetas'][:n_iter]) is None
    for c in (1, 2):
        ref_trace, ref_ops = single_chain(method, X, y, N, seeds[c], n_iter)
        assert first_divergence(out['thetas'][c], ref_trace) is None
        assert out['n_cubic_ops'][c] == ref_ops
    assert out['rounds'] >= n_iter


@pytest.mark.parametrize('method', ['ess+rdss', 'pmmh'])
def test_chain_groups_do_not_change_the_chains(method):
    """A list of backends = chain groups on their own scheduler threads: every chain's trace, reject counts and
    cubic-op counts equal those of the single-group run."""
    g = load_golden('samplers')
    X, y = g['X'], g['y']
    N, n_iter = 4, 60
    seeds = [1000 + N, 77, 4242, 5, 6]
    mk = lambda backends: batched.BatchedAPMSampler(backends, X.shape[0], N, 2, method, batched.make_log_prior(X.shape[1], False),
                                                    seeds, prop_scales=[0.5, 0.5], slice_width=1.)
    init = lambda prng: synth.draw_theta_prior(prng, X.shape[1], ard=False)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        one = mk(OracleBackend(X, y)).get_samples(None, n_iter, theta_init_sampler=init)
        two = mk([OracleBackend(X, y), OracleBackend(X, y)]).get_samples(None, n_iter, theta_init_sampler=init)
    assert np.all(two['failed'] == 0)
    for k in ('thetas', 'n_reject', 'n_cubic_ops', 'n_full', 'n_cached'):
        assert np.array_equal(one[k], two[k]), k
    assert first_divergence(two['thetas'][0], g['%s_N%d_thetas' % (method, N)][:n_iter]) is None


def test_shard_chains_partition():
    from apm_b200.distributed import shard_chains
    for n, w in [(10, 1), (10, 3), (256, 8), (5, 8)]:
        parts = [shard_chains(n, r, w) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
