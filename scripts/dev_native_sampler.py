"""Native sampler (apm_sampler_run) at the headline shape: chain-iterations/s and scheduling statistics.
env: B (chains), ITERS, METHOD, FRACS, JOBS, MIN2 (-> APM_SAMPLER_BATCH_FRAC / _JOBS / _MIN_SECOND, read by apm_sampler_create)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apm_b200 import _capi, batched, synth
n, D, N = int(os.environ.get('NDATA', 768)), 8, 64
method = os.environ.get('METHOD', 'ess+rdss')
iters = int(os.environ.get('ITERS', 60))
X, y, th = synth.make_dataset(n, D, seed=0)
for B in [int(b) for b in os.environ.get('B', '256').split(',')]:
    eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=2 * B, max_nimp=N)
    cfgs = [(float(f), int(j), int(m2)) for f in os.environ.get('FRACS', '0.5').split(',') for j in os.environ.get('JOBS', '1').split(',')
            for m2 in (os.environ.get('MIN2', '0').split(',') if int(j) > 1 else ['0'])]
    for frac, jobs, min2 in cfgs:
        os.environ['APM_SAMPLER_BATCH_FRAC'] = str(frac)
        os.environ['APM_SAMPLER_JOBS'] = str(jobs)
        os.environ['APM_SAMPLER_MIN_SECOND'] = str(min2)
        drv = batched.BatchedAPMSampler(batched.EngineBackend(eng), n, N, D + 1, method, batched.make_log_prior(D, True),
                                        [1000 + c for c in range(B)], prop_scales=np.full(D + 1, 0.1), rng='native')
        th0 = synth.bulk_thetas(B, D, seed=1000)
        drv.get_samples(th0, 3)
        t0 = time.perf_counter()
        out = drv.get_samples(th0, iters + 1)
        dt = time.perf_counter() - t0
        s = drv.async_stats
        print('%s B=%d frac=%.2f jobs=%d min2=%d iters=%d: %.0f chain-iters/s | FULL calls %d, %.1f chains/call, %.2f ms/call, >=1 in flight %.0f%% of the run, '
              '%.0f FULL est/s while busy | CACHED calls %d, %.1f chains/call | full/iter %.2f cached/iter %.2f failed %d'
              % (method, B, frac, jobs, min2, iters, B * iters / dt, s['full_calls'], s['full_chains'] / max(s['full_calls'], 1),
                 1e3 * s['t_flight'] / max(s['full_calls'], 1), 100 * s['t_busy'] / s['t_total'], s['full_chains'] / s['t_busy'],
                 s['cached_calls'], s['cached_chains'] / max(s['cached_calls'], 1), (out['n_full'].mean() - 1) / iters,
                 out['n_cached'].mean() / iters, int((out['failed'] != 0).sum())), flush=True)
        drv._native.close()
    eng.close()
