"""FULL-estimate throughput of the bench workload under the environment's lane settings (APM_LANES, ...)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apm_b200 import _capi, synth
if os.environ.get('LIB'):
    _capi.LIB_PATH = os.path.join(ROOT, os.environ['LIB'])
import torch
n, D, N, B = 768, 8, 64, int(os.environ.get('B', 256))
reps = int(os.environ.get('REPS', 12))
X, y, th = synth.make_dataset(n, D, seed=0)
eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=B, max_nimp=N)
if os.environ.get('MAXIT'):
    eng.set_newton(1e-4, int(os.environ['MAXIT']))   # experiment: retire the stragglers (they fail with status 2)
thetas = [synth.bulk_thetas(B, D, seed=s) for s in range(4)]
us = [torch.randn(B, n, N, dtype=torch.float64, device='cuda') for _ in range(2)]
slots = np.arange(B)
for i in range(3):
    eng.estimate_full(thetas[i % 4], us[i % 2], slots)
torch.cuda.synchronize(); t = time.perf_counter()
for i in range(reps):
    out = eng.estimate_full(thetas[i % 4], us[i % 2], slots)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / reps
print('%-40s FULL %.2f ms -> %.0f est/s (iters mean %.2f)' % (os.environ.get('TAG', ''), dt * 1e3, B / dt, (out[1] - 3).mean()), flush=True)
