"""Long native-sampler runs at the headline shape: no failed chains, finite traces, plausible acceptance, stable rate."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apm_b200 import _capi, batched, synth, utils
n, D, N = 768, 8, 64
X, y, th = synth.make_dataset(n, D, seed=0)
for method, B, iters in (('ess+rdss', 256, int(os.environ.get('ITERS', 600))), ('pmmh', 512, 300), ('mi+mh', 256, 300)):
    eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=2 * B, max_nimp=N)
    drv = batched.BatchedAPMSampler(batched.EngineBackend(eng), n, N, D + 1, method, batched.make_log_prior(D, True),
                                    [5000 + c for c in range(B)], prop_scales=np.full(D + 1, 0.1), rng='native')
    th0 = synth.bulk_thetas(B, D, seed=5000)
    t0 = time.perf_counter()
    out = drv.get_samples(th0, iters + 1)
    dt = time.perf_counter() - t0
    tr = out['thetas']
    half = tr[:, iters // 2:, :]
    rhat = max(utils.gelman_rubin(half[:64, :, k]) for k in range(D + 1))
    print('%s: %d chains x %d iterations in %.1f s = %.0f it/s; failed %d; finite %s; reject rates u %.2f theta %.2f; '
          'posterior mean theta[0] %.3f (sd over chains %.3f); max R-hat over components (64 chains, 2nd half) %.2f; full/iter %.2f'
          % (method, B, iters, dt, B * iters / dt, int((out['failed'] != 0).sum()), bool(np.all(np.isfinite(tr))),
             out['n_reject'][:, 0].mean() / iters, out['n_reject'][:, 1].mean() / iters, half[:, :, 0].mean(),
             half[:, :, 0].mean(axis=1).std(), rhat, (out['n_full'].mean() - 1) / iters), flush=True)
    drv._native.close()
    eng.close()
