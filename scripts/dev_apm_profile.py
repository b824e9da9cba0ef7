import os, sys, time, cProfile, pstats
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from apm_b200 import _capi, synth, batched
n, D, N, B = 768, 8, 64, 256
X, y, th = synth.make_dataset(n, D, seed=0)
eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=2 * B, max_nimp=N)
eng.use_torch_stream()
dev = torch.device('cuda', 0)
drv = batched.BatchedAPMSampler(batched.EngineBackend(eng), n, N, D + 1, 'ess+rdss', batched.make_log_prior(D, True),
                                [1000 + c for c in range(B)], rng='device', device=dev)
thetas = synth.bulk_thetas(B, D)
drv.get_samples(thetas, 3)
torch.cuda.synchronize()
pr = cProfile.Profile()
t = time.time()
pr.enable()
out = drv.get_samples(thetas, 11)
pr.disable()
torch.cuda.synchronize()
dt = time.time() - t
print('10 iterations: %.3f s, rounds %d, full/iter %.2f cached/iter %.2f' % (dt, out['rounds'], (out['n_full'].mean() - 1) / 10, out['n_cached'].mean() / 10))
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
