"""Where does a batched ESS+RD-SS iteration spend its time (calls, batch sizes, host overhead)?"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from apm_b200 import _capi, batched, synth
n, D, N, B = 768, 8, 64, int(os.environ.get('B', 256))
method = os.environ.get('METHOD', 'ess+rdss')
iters = int(os.environ.get('ITERS', 10))
X, y, th = synth.make_dataset(n, D, seed=0)
G = int(os.environ.get('CHAIN_GROUPS', 1))
if G > 1:      # chain groups: one context and one scheduler thread per group, no per-call timing
    per = (B + G - 1) // G
    engs = [_capi.Engine(X, y, kernel='ard', max_chains=per, n_slots=2 * per, max_nimp=N) for _ in range(G)]
    dev = torch.device('cuda', 0)
    drv = batched.BatchedAPMSampler([batched.EngineBackend(e) for e in engs], n, N, D + 1, method, batched.make_log_prior(D, True),
                                    [1000 + c for c in range(B)], prop_scales=np.full(D + 1, 0.1), rng='device', device=dev,
                                    full_batch_frac=float(os.environ.get('FRAC', 0.8)), async_full=bool(int(os.environ.get('ASYNC', 0))),
                                    async_batch_frac=float(os.environ.get('AFRAC', 0.5)))
    th0 = synth.bulk_thetas(B, D, seed=1000)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = drv.get_samples(th0, iters + 1)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print('%s B=%d groups=%d iters=%d: %.3f s -> %.0f chain-iters/s; rounds %d; full/iter %.2f cached/iter %.2f failed %d'
              % (method, B, G, iters, dt, B * iters / dt, out['rounds'], (out['n_full'].mean() - 1) / iters,
                 out['n_cached'].mean() / iters, int((out['failed'] != 0).sum())), flush=True)
    sys.exit(0)
eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=2 * B, max_nimp=N)
eng.use_torch_stream()
log = {'full': [], 'cached': []}
of, oc = eng.estimate_full, eng.estimate_cached
def tf(*a):
    torch.cuda.synchronize(); t = time.perf_counter(); r = of(*a); log['full'].append((len(r[0]), time.perf_counter() - t)); return r
def tc(*a):
    torch.cuda.synchronize(); t = time.perf_counter(); r = oc(*a); log['cached'].append((len(r[0]), time.perf_counter() - t)); return r
if not int(os.environ.get('ASYNC', 0)):      # per-call timing synchronises the device: not with the asynchronous scheduler
    eng.estimate_full, eng.estimate_cached = tf, tc
dev = torch.device('cuda', 0)
drv = batched.BatchedAPMSampler(batched.EngineBackend(eng), n, N, D + 1, method, batched.make_log_prior(D, True),
                                [1000 + c for c in range(B)], prop_scales=np.full(D + 1, 0.1), rng='device', device=dev,
                                full_batch_frac=float(os.environ.get('FRAC', 0.8)), async_full=bool(int(os.environ.get('ASYNC', 0))),
                                async_batch_frac=float(os.environ.get('AFRAC', 0.5)))
th0 = synth.bulk_thetas(B, D, seed=1000)
for rep in range(2):
    log['full'].clear(); log['cached'].clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = drv.get_samples(th0, iters + 1)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    tf_, tc_ = sum(t for _, t in log['full']), sum(t for _, t in log['cached'])
    print('%s B=%d iters=%d: %.3f s -> %.0f chain-iters/s; rounds %d; FULL calls %d (%.3f s) sizes %s; CACHED calls %d (%.3f s) sizes %s; other %.3f s'
          % (method, B, iters, dt, B * iters / dt, out['rounds'], len(log['full']), tf_, [b for b, _ in log['full']][:40],
             len(log['cached']), tc_, [b for b, _ in log['cached']][:40], dt - tf_ - tc_), flush=True)
    if getattr(drv, 'async_stats', None):
        print('   async:', {k: (round(v, 3) if isinstance(v, float) else v) for k, v in drv.async_stats.items()}, flush=True)
    if os.environ.get('DUMP'):
        print('FULL (size, ms):', ' '.join('%d:%.1f' % (b, t * 1e3) for b, t in log['full']))
        print('CACHED (size, ms):', ' '.join('%d:%.2f' % (b, t * 1e3) for b, t in log['cached']))
