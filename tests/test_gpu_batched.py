"""Batched lock-step sampler driver on the GPU: chain 0 of a batch reproduces the golden chain of the
unmodified reference (accept/reject sequence over its first iterations); device-resident RNG mode runs."""
import warnings

import numpy as np
import pytest

from apm_b200 import _capi, batched, synth
from conftest import load_golden
from wiring import first_divergence

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('method', ['mi+mh', 'ess+mh', 'mi+rdss', 'ess+rdss', 'pmmh'])
def test_batched_gpu_matches_reference_chain(method):
    g = load_golden('samplers')
    X, y = g['X'], g['y']
    N, n_iter = 4, 300
    seeds = [1000 + N] + [50 + c for c in range(7)]
    B = len(seeds)
    eng = _capi.Engine(X, y, kernel='iso', max_chains=B, n_slots=2 * B, max_nimp=N)
    drv = batched.BatchedAPMSampler(batched.EngineBackend(eng), X.shape[0], N, 2, method,
                                    batched.make_log_prior(X.shape[1], False), seeds, prop_scales=[0.5, 0.5])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        out = drv.get_samples(None, n_iter, theta_init_sampler=lambda prng: synth.draw_theta_prior(prng, X.shape[1], ard=False))
    assert np.all(out['failed'] == 0)
    key = '%s_N%d_' % (method, N)
    div = first_divergence(out['thetas'][0], g[key + 'thetas'][:n_iter])
    assert div is None, 'batched chain diverges from the reference at iteration %d' % div
    assert np.all(np.isfinite(out['thetas']))
    eng.close()


def test_batched_device_rng_mode():
    import torch
    X, y, th = synth.make_dataset(96, 3, seed=2)
    B, N = 6, 8
    eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=2 * B, max_nimp=N)
    eng.use_torch_stream()
    drv = batched.BatchedAPMSampler(batched.EngineBackend(eng), 96, N, 4, 'ess+rdss', batched.make_log_prior(3, True),
                                    [10 + c for c in range(B)], rng='device', device=torch.device('cuda', 0))
    out = drv.get_samples(np.tile(th, (B, 1)), 20)
    assert np.all(out['failed'] == 0) and np.all(np.isfinite(out['thetas']))
    assert np.all(out['n_full'] >= 20) and np.all(out['n_cached'] >= 19)
    # chains are distinct and moved
    assert np.std(out['thetas'][:, -1, 0]) > 0
    eng.close()


@pytest.mark.parametrize('method', ['ess+rdss', 'mi+mh', 'pmmh'])
def test_batched_async_full_scheduler(method):
    """Device-RNG mode with asynchronous FULL estimates: the FULL call runs on a worker thread / stream while the
    scheduler serves CACHED requests through a companion context (shared cache slots).  The chains must behave like
    the synchronous scheduler's: every cached estimate equals a fresh FULL estimate for the same (theta, u), which the
    chain checks itself through its log_f bookkeeping -- here: finite traces, plausible acceptance, no failures, and the
    same posterior location as the synchronous run."""
    import torch
    X, y, th = synth.make_dataset(96, 3, seed=2)
    B, N, iters = 24, 8, 60
    dev = torch.device('cuda', 0)
    res = {}
    for mode in (False, True):
        eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=2 * B, max_nimp=N)
        eng.use_torch_stream()
        drv = batched.BatchedAPMSampler(batched.EngineBackend(eng), 96, N, 4, method, batched.make_log_prior(3, True),
                                        [10 + c for c in range(B)], prop_scales=np.full(4, 0.3), rng='device', device=dev,
                                        async_full=mode)
        out = drv.get_samples(np.tile(th, (B, 1)), iters)
        assert np.all(out['failed'] == 0) and np.all(np.isfinite(out['thetas']))
        assert np.all(out['n_full'] >= iters - 1)
        res[mode] = out
        eng.close()
    # same target distribution: pooled second-half means agree within Monte-Carlo error
    a = res[False]['thetas'][:, iters // 2:].reshape(-1, 4)
    b = res[True]['thetas'][:, iters // 2:].reshape(-1, 4)
    assert np.all(np.abs(a.mean(0) - b.mean(0)) < 4. * (a.std(0) + b.std(0)) / np.sqrt(B))


def test_companion_context_shares_slots_and_refuses_full():
    X, y, th = synth.make_dataset(70, 2, seed=4)
    rs = np.random.RandomState(0)
    B, N = 3, 5
    eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=2 * B, max_nimp=N)
    comp = eng.companion()
    thetas = th[None] + 0.1 * rs.normal(size=(B, 3))
    u = rs.normal(size=(B, 70, N))
    full, _, st = eng.estimate_full(thetas, u, [0, 2, 4])
    assert np.all(st == 0)
    c_own, _ = eng.estimate_cached([0, 2, 4], u)
    c_comp, _ = comp.estimate_cached([0, 2, 4], u)
    assert np.array_equal(c_own, c_comp) and np.allclose(c_comp, full, rtol=1e-12)
    with pytest.raises(_capi.ApmError):
        comp.estimate_full(thetas, u, [1, 3, 5])
    with pytest.raises(_capi.ApmError):
        comp.estimate_cached([1], u[:1])          # slot 1 holds no cache: validity flags are the owner's
    eng.close()


def test_chain_groups_on_two_contexts_match_single_group():
    """Two engine contexts driven by two scheduler threads (chain groups): identical per-chain traces in parity mode,
    chain 0 still reproduces the reference's golden chain; device-RNG mode runs on per-group streams."""
    import torch
    g = load_golden('samplers')
    X, y = g['X'], g['y']
    N, n_iter, method = 4, 150, 'ess+rdss'
    seeds = [1000 + N] + [50 + c for c in range(7)]
    B = len(seeds)
    engs = [_capi.Engine(X, y, kernel='iso', max_chains=B, n_slots=2 * B, max_nimp=N) for _ in range(3)]
    mk = lambda backends, **kw: batched.BatchedAPMSampler(backends, X.shape[0], N, 2, method, batched.make_log_prior(X.shape[1], False),
                                                          seeds, prop_scales=[0.5, 0.5], **kw)
    init = lambda prng: synth.draw_theta_prior(prng, X.shape[1], ard=False)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        one = mk(batched.EngineBackend(engs[0])).get_samples(None, n_iter, theta_init_sampler=init)
        two = mk([batched.EngineBackend(engs[1]), batched.EngineBackend(engs[2])]).get_samples(None, n_iter, theta_init_sampler=init)
    assert np.all(two['failed'] == 0)
    for k in ('thetas', 'n_reject', 'n_cubic_ops', 'n_full', 'n_cached'):
        assert np.array_equal(one[k], two[k]), k
    assert first_divergence(two['thetas'][0], g['%s_N%d_thetas' % (method, N)][:n_iter]) is None
    dev = torch.device('cuda', 0)
    drv = mk([batched.EngineBackend(engs[1]), batched.EngineBackend(engs[2])], rng='device', device=dev)
    out = drv.get_samples(np.tile(g['%s_N%d_thetas' % (method, N)][0], (B, 1)), 20)
    assert np.all(out['failed'] == 0) and np.all(np.isfinite(out['thetas']))
    assert np.std(out['thetas'][:, -1, 0]) > 0
    for e in engs:
        e.close()


@pytest.mark.parametrize('tag,method', [('pima', 'ess+rdss'), ('pima', 'mi+mh'), ('pima', 'pmmh'), ('breast', 'ess+rdss')])
def test_headline_shape_chains_match_reference(tag, method):
    """Accept/reject parity where the benchmark is quoted (BASELINE configs 1-2): chains of the UNMODIFIED reference at
    pima (n = 768, D = 8) and breast (n = 682, D = 9) shape with N_imp = 64 (tests/golden/samplers_fullsize.npz,
    oracle/gen_golden.py:gen_samplers_fullsize) against chain 0 of a lock-step batch on the GPU: identical theta traces,
    reject counts and cubic-op counts over all golden iterations."""
    g = load_golden('samplers_fullsize')
    X, y = g[tag + '_X'], g[tag + '_y']
    N = 64
    key = '%s_%s_N%d_' % (tag, method, N)
    ref = g[key + 'thetas']
    n_iter = ref.shape[0]
    seeds = [1000 + N, 7, 8, 9]
    B = len(seeds)
    eng = _capi.Engine(X, y, kernel='iso', max_chains=B, n_slots=2 * B, max_nimp=N)
    drv = batched.BatchedAPMSampler(batched.EngineBackend(eng), X.shape[0], N, 2, method,
                                    batched.make_log_prior(X.shape[1], False), seeds, prop_scales=[0.5, 0.5])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        out = drv.get_samples(None, n_iter, theta_init_sampler=lambda prng: synth.draw_theta_prior(prng, X.shape[1], ard=False))
    assert np.all(out['failed'] == 0)
    div = first_divergence(out['thetas'][0], ref)
    assert div is None, 'chain diverges from the reference at iteration %d' % div
    assert out['n_cubic_ops'][0] == int(g[key + 'cubic_ops'])
    if g[key + 'n_reject'].size:
        assert np.array_equal(out['n_reject'][0][-g[key + 'n_reject'].size:], g[key + 'n_reject'])
    eng.close()


@pytest.mark.parametrize('method', ['ess+rdss', 'mi+mh'])
def test_headline_shape_1000_iterations_match_reference(method):
    """The north-star's bar at the shape the benchmark is quoted on: the accept/reject sequence of the UNMODIFIED reference
    over the first 1000 iterations at pima shape (n = 768, D = 8, N_imp = 64; tests/golden/samplers_long_*.npz,
    oracle/gen_golden.py:gen_samplers_long) against chain 0 of a lock-step batch on the GPU: zero divergence, same reject and
    cubic-op counts."""
    g = load_golden('samplers_long_' + method.replace('+', '_'))
    n, D, N, n_iter = int(g['n']), int(g['D']), int(g['N']), int(g['n_iter'])
    X, y, _ = synth.make_dataset(n, D, seed=int(g['data_seed']))
    ref = g['thetas']
    assert ref.shape[0] == n_iter == 1000
    seeds = [int(g['chain_seed']), 7, 8, 9]
    B = len(seeds)
    eng = _capi.Engine(X, y, kernel='iso', max_chains=B, n_slots=2 * B, max_nimp=N)
    drv = batched.BatchedAPMSampler(batched.EngineBackend(eng), n, N, 2, method,
                                    batched.make_log_prior(D, False), seeds, prop_scales=[0.5, 0.5])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        out = drv.get_samples(None, n_iter, theta_init_sampler=lambda prng: synth.draw_theta_prior(prng, D, ard=False))
    assert np.all(out['failed'] == 0)
    div = first_divergence(out['thetas'][0], ref)
    assert div is None, 'chain diverges from the reference at iteration %d of 1000' % div
    assert out['n_cubic_ops'][0] == int(g['cubic_ops'])
    if g['n_reject'].size:
        assert np.array_equal(out['n_reject'][0][-g['n_reject'].size:], g['n_reject'])
    eng.close()
