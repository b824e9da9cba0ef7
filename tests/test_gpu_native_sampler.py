"""Native batched sampler (apm_sampler_* of the C ABI, csrc/sampler.cuh) against the Python chain generators of
apm_b200.batched -- which are pinned to the reference's chains by tests/test_gpu_batched.py -- on the SAME Philox streams
(rng='philox': numpy mirror of the device generator, u drawn on the host and uploaded in the reference layout)."""
import numpy as np
import pytest

from apm_b200 import _capi, batched, synth
from conftest import load_golden

pytestmark = pytest.mark.gpu


def _run(X, y, kernel, n_theta, method, seeds, N, theta_init, iters, rng, prop_scales=None, slice_width=1.):
    B = len(seeds)
    eng = _capi.Engine(X, y, kernel=kernel, max_chains=B, n_slots=2 * B, max_nimp=N)
    drv = batched.BatchedAPMSampler(batched.EngineBackend(eng), X.shape[0], N, n_theta, method,
                                    batched.make_log_prior(X.shape[1], kernel == 'ard'), seeds, prop_scales=prop_scales,
                                    slice_width=slice_width, rng=rng)
    out = drv.get_samples(theta_init, iters)
    stats = getattr(drv, 'async_stats', None)
    if rng == 'native':
        drv._native.close()
    eng.close()
    return out, stats


@pytest.mark.parametrize('method', ['mi+mh', 'ess+mh', 'mi+rdss', 'ess+rdss', 'pmmh'])
def test_native_sampler_matches_python_chain_logic(method):
    g = load_golden('samplers')
    X, y = g['X'], g['y']
    N, iters = 4, 120
    seeds = [2000 + 7 * c for c in range(6)]
    rs = np.random.RandomState(5)
    theta_init = np.stack([synth.draw_theta_prior(rs, X.shape[1], ard=False) for _ in seeds])
    ref, _ = _run(X, y, 'iso', 2, method, seeds, N, theta_init, iters, 'philox', prop_scales=[0.5, 0.5])
    out, stats = _run(X, y, 'iso', 2, method, seeds, N, theta_init, iters, 'native', prop_scales=[0.5, 0.5])
    assert np.all(ref['failed'] == 0) and np.all(out['failed'] == 0)
    # identical accept / reject sequences (same counters) and the same traces: the only differences are device vs libm
    # log / sincos in the normals (~1e-16) in front of the same estimator
    assert np.array_equal(out['n_reject'], ref['n_reject'])
    assert np.array_equal(out['n_full'], ref['n_full']) and np.array_equal(out['n_cached'], ref['n_cached'])
    assert np.array_equal(out['n_cubic_ops'], ref['n_cubic_ops'])
    np.testing.assert_allclose(out['thetas'], ref['thetas'], rtol=1e-9, atol=1e-12)
    assert stats['full_chains'] == int(out['n_full'].sum())
    assert stats['cached_chains'] == int(out['n_cached'].sum())


def test_native_sampler_odd_sample_count_and_ard():
    X, y, th = synth.make_dataset(96, 3, seed=2)
    seeds = [11, 12, 13, 14, 15]
    theta_init = np.tile(th, (len(seeds), 1)) + 0.1 * np.random.RandomState(0).normal(size=(len(seeds), 4))
    for N in (1, 3):
        ref, _ = _run(X, y, 'ard', 4, 'ess+rdss', seeds, N, theta_init, 40, 'philox')
        out, _ = _run(X, y, 'ard', 4, 'ess+rdss', seeds, N, theta_init, 40, 'native')
        assert np.all(out['failed'] == 0)
        assert np.array_equal(out['n_full'], ref['n_full']) and np.array_equal(out['n_cached'], ref['n_cached'])
        np.testing.assert_allclose(out['thetas'], ref['thetas'], rtol=1e-9, atol=1e-12)


def test_native_sampler_trace_depends_on_the_seed_only():
    """A chain's trace is a function of its seed and start: running it in another batch (other batch-mates, another
    position, other FULL / CACHED call compositions) gives bit-identical states."""
    X, y, th = synth.make_dataset(96, 3, seed=2)
    seeds = [100 + c for c in range(12)]
    theta_init = np.tile(th, (12, 1)) + 0.2 * np.random.RandomState(1).normal(size=(12, 4))
    full, _ = _run(X, y, 'ard', 4, 'ess+rdss', seeds, 8, theta_init, 30, 'native')
    part, _ = _run(X, y, 'ard', 4, 'ess+rdss', seeds[4:9], 8, theta_init[4:9], 30, 'native')
    assert np.array_equal(full['thetas'][4:9], part['thetas'])
    assert np.array_equal(full['n_cubic_ops'][4:9], part['n_cubic_ops'])
    # chains are distinct and moved
    assert np.std(full['thetas'][:, -1, 0]) > 0


def test_native_sampler_two_full_calls_in_flight_give_the_same_chains(monkeypatch):
    """APM_SAMPLER_JOBS=2: a second FULL call (companion context with full workspaces on the same cache slots) may start
    while one is in flight.  Scheduling only: every chain's trace, counters and operation counts are unchanged."""
    X, y, th = synth.make_dataset(96, 3, seed=2)
    seeds = [300 + c for c in range(20)]
    theta_init = np.tile(th, (20, 1)) + 0.2 * np.random.RandomState(3).normal(size=(20, 4))
    one, s1 = _run(X, y, 'ard', 4, 'ess+rdss', seeds, 8, theta_init, 40, 'native')
    monkeypatch.setenv('APM_SAMPLER_JOBS', '2')
    monkeypatch.setenv('APM_SAMPLER_MIN_SECOND', '2')
    two, s2 = _run(X, y, 'ard', 4, 'ess+rdss', seeds, 8, theta_init, 40, 'native')
    assert np.all(two['failed'] == 0)
    assert np.array_equal(one['thetas'], two['thetas'])
    for k in ('n_reject', 'n_full', 'n_cached', 'n_cubic_ops'):
        assert np.array_equal(one[k], two[k]), k
    assert s2['full_chains'] == s1['full_chains'] and s2['full_calls'] >= s1['full_calls']


def test_native_sampler_headline_shape_runs():
    X, y, th = synth.make_dataset(768, 8, seed=0)
    B, N = 48, 64
    thetas = synth.bulk_thetas(B, 8, seed=0)
    out, stats = _run(X, y, 'ard', 9, 'ess+rdss', [7 + c for c in range(B)], N, thetas, 6, 'native')
    assert np.all(out['failed'] == 0) and np.all(np.isfinite(out['thetas']))
    assert np.all(out['n_full'] >= 6) and np.all(out['n_cached'] >= 5)
    assert stats['full_calls'] >= 6
