"""Histogram of Newton iteration counts over the bench workload's chains (lock-step rounds = max)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from apm_b200 import _capi, synth
n, D, N, B = 768, 8, 64, 256
X, y, th = synth.make_dataset(n, D, seed=0)
eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=B, max_nimp=N)
u = torch.randn(B, n, N, dtype=torch.float64, device='cuda')
for seed in (1234, 1251, 1268):
    thetas = synth.bulk_thetas(B, D, seed=seed)
    out = eng.estimate_full(thetas, u, np.arange(B))
    it = out[1] - 3
    print('seed', seed, 'iters histogram', np.bincount(it), 'mean %.3f' % it.mean(), flush=True)
