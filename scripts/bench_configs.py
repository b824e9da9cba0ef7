#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations on ONE GPU (bench.py stays the headline: config 1/pima FULL
estimates).  One JSON line per configuration; chains are independent, so the N-GPU figure of a sharded configuration
is N x the per-GPU figure at the per-GPU chain count used here (bench.py --gpus N measures that scaling).

    python scripts/bench_configs.py [breast] [nimp] [ep] [pmmh] [large]

  breast  config 2: breast-shaped synthetic (n=682, D=9), Laplace IS N_imp=64, E-SS-u + RD-SS-theta, 256 lock-step chains
  nimp    config 3: pima-shaped, N_imp sweep 1..1024 with the Laplace approximation (the reference has no EP: SURVEY App. D),
          256 chains per GPU (= 1024 chains over 4 GPUs)
  ep      config 3 as named: the same workload at N_imp=64 with the EP approximation (extension)
  pmmh    config 4: pseudo-marginal MH (fresh u inside every estimate), 512 chains per GPU (= 4096 chains over 8 GPUs)
  large   config 5: n=8192, D=16 ARD, E-SS-u + RD-SS-theta, 8 chains per GPU
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from apm_b200 import _capi, batched, synth  # noqa: E402

DEV = torch.device('cuda', 0)


def timed_ms(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    outs = [fn(i) for i in range(reps)]
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, outs


def estimator_rates(eng, n, D, N, B, reps=5, seed=3):
    """FULL and CACHED estimates/s with device-resident u (events on the launching stream)."""
    gen = torch.Generator(device=DEV)
    gen.manual_seed(seed)
    u = [torch.randn(B, n, N, dtype=torch.float64, device=DEV, generator=gen) for _ in range(2)]
    thetas = [synth.bulk_thetas(B, D, seed=seed + 11 * i) for i in range(2)]
    slots = [np.arange(B), np.arange(B, 2 * B)]
    for i in range(2):
        eng.estimate_full(thetas[i], u[i], slots[i])
    eng.work_count(reset=True)
    ms_full, outs = timed_ms(lambda i: eng.estimate_full(thetas[i % 2], u[i % 2], slots[i % 2]), reps)
    estimator_rates.units_per_estimate = sum(eng.work_count(reset=True)) / float(B * reps)   # executed n^3/3 units (chol + M' builds)
    ms_cached, _ = timed_ms(lambda i: eng.estimate_cached(slots[i % 2], u[(i + 1) % 2]), reps)
    iters = float(np.mean([(o[1] - 3).mean() for o in outs]))
    bad = int(sum((o[2] != 0).sum() for o in outs))
    return B / ms_full * 1e3, B / ms_cached * 1e3, iters, bad


def sampler_rate(eng, n, D, N, B, method, iters, seed=1000):
    drv = batched.BatchedAPMSampler(batched.EngineBackend(eng), n, N, D + 1, method, batched.make_log_prior(D, True),
                                    [seed + c for c in range(B)], prop_scales=np.full(D + 1, 0.1), rng='device', device=DEV,
                                    async_full=True)
    th0 = synth.bulk_thetas(B, D, seed=seed)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = drv.get_samples(th0, iters + 1)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {'value': B * iters / dt, 'unit': 'chain-iterations/s', 'iterations': iters, 'chains': B, 'method': method,
            'full_estimates_per_iter': float(out['n_full'].mean() - 1) / iters,
            'cached_estimates_per_iter': float(out['n_cached'].mean()) / iters,
            'failed_chains': int((out['failed'] != 0).sum()),
            'timing': 'host wall clock incl. the Python scheduler and the drain of the last iterations, device RNG, asynchronous FULL rounds'}


def line(name, workload, metric, value, unit, extra):
    d = {'config': name, 'metric': metric, 'value': value, 'unit': unit, 'n_gpus': 1, 'dtype': 'f64', 'data': 'synthetic',
         'workload': workload}
    d.update(extra)
    print(json.dumps(d), flush=True)


def run_breast():
    n, D, N, B = 682, 9, 64, 256
    X, y, _ = synth.make_dataset(n, D, seed=0)
    eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=2 * B, max_nimp=N)
    eng.use_torch_stream()
    full, cached, iters, bad = estimator_rates(eng, n, D, N, B)
    apm = sampler_rate(eng, n, D, N, B, 'ess+rdss', 60)
    line('config 2 (breast)', 'breast-shaped synthetic (n=682, D=9), Laplace IS N_imp=64, 256 chains on 1 GPU',
         'APM-MCMC iterations/s (E-SS-u + RD-SS-theta)', apm['value'], apm['unit'],
         {'apm': apm, 'full_estimates_per_s': full, 'cached_estimates_per_s': cached, 'newton_iters_mean': iters,
          'failed_chains': bad})
    eng.close()


def run_nimp():
    n, D, B = 768, 8, 256
    X, y, _ = synth.make_dataset(n, D, seed=0)
    sweep = {}
    for N in (1, 4, 16, 64, 256, 1024):
        eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=2 * B, max_nimp=N)
        eng.use_torch_stream()
        full, cached, iters, bad = estimator_rates(eng, n, D, N, B, reps=4)
        sweep[str(N)] = {'full_estimates_per_s': full, 'cached_estimates_per_s': cached, 'failed_chains': bad}
        eng.close()
    line('config 3 (N_imp sweep)', 'pima-shaped synthetic (n=768, D=8), Laplace IS (the reference has no EP), 256 chains per GPU',
         'FULL log-ML estimates/s at N_imp=64', sweep['64']['full_estimates_per_s'], 'estimates/s', {'n_imp_sweep': sweep})


def run_ep():
    n, D, N, B = 768, 8, 64, 256
    X, y, _ = synth.make_dataset(n, D, seed=0)
    eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=2 * B, max_nimp=N)
    eng.use_torch_stream()
    eng.set_approximation('ep', 1e-6, 100, 1.0)
    full, cached, iters, bad = estimator_rates(eng, n, D, N, B, reps=3)
    line('config 3 (EP extension)', 'pima-shaped synthetic (n=768, D=8), EP posterior approximation (extension: not in the reference; '
         'checked against the parallel-EP restatement in oracle/), IS N_imp=64, 256 chains per GPU',
         'FULL log-ML estimates/s with the EP approximation', full, 'estimates/s',
         {'cached_estimates_per_s': cached, 'ep_iters_mean': iters, 'failed_chains': bad, 'ep_tol': 1e-6, 'ep_damping': 1.0})
    eng.close()


def run_pmmh():
    n, D, N, B = 768, 8, 64, 512
    X, y, _ = synth.make_dataset(n, D, seed=0)
    eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=2 * B, max_nimp=N)
    eng.use_torch_stream()
    apm = sampler_rate(eng, n, D, N, B, 'pmmh', 30)
    line('config 4 (PM-MH)', 'pima-shaped synthetic, pseudo-marginal MH (fresh u in every estimate), 512 chains per GPU',
         'PM-MH iterations/s', apm['value'], apm['unit'], {'apm': apm})
    eng.close()


def run_large():
    n, D, N, B = 8192, 16, 64, 8
    X, y, _ = synth.make_dataset(n, D, seed=0)
    eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=2 * B, max_nimp=N)
    eng.use_torch_stream()
    full, cached, iters, bad = estimator_rates(eng, n, D, N, B, reps=2)
    units = estimator_rates.units_per_estimate
    apm = sampler_rate(eng, n, D, N, B, 'ess+rdss', 2)
    n3 = float(n)**3
    line('config 5 (large)', 'large synthetic GP probit (n=8192, D=16, ARD), Laplace IS N_imp=64, 8 chains per GPU',
         'FULL log-ML estimates/s', full, 'estimates/s',
         {'cached_estimates_per_s': cached, 'newton_iters_mean': iters, 'failed_chains': bad, 'apm': apm,
          'survey_tflops': full * (iters / 3. + 8. / 3.) * n3 / 1e12,
          'executed_tflops': full * units / 3. * n3 / 1e12, 'executed_n3_over_3_units_per_estimate': units})
    eng.close()


if __name__ == '__main__':
    which = sys.argv[1:] or ['breast', 'nimp', 'ep', 'pmmh', 'large']
    for w in which:
        {'breast': run_breast, 'nimp': run_nimp, 'ep': run_ep, 'pmmh': run_pmmh, 'large': run_large}[w]()
