"""Repeated FULL + CACHED estimates of the headline batch (256 chains, n = 768, N_imp = 64): identical bits every time."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from apm_b200 import _capi, synth
n, D, N, B = 768, 8, 64, int(os.environ.get('B', 256))
X, y, th = synth.make_dataset(n, D, seed=0)
eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=B, max_nimp=N)
thetas = synth.bulk_thetas(B, D, seed=3)
u = torch.randn(B, n, N, dtype=torch.float64, device='cuda')
u2 = torch.randn(B, n, N, dtype=torch.float64, device='cuda')
ref = None
for rep in range(int(os.environ.get('REPS', 12))):
    val, ops, st = eng.estimate_full(thetas, u, np.arange(B))
    cval, _ = eng.estimate_cached(np.arange(B), u2)
    if ref is None:
        ref = (val.copy(), ops.copy(), cval.copy())
    same = np.array_equal(val, ref[0]) and np.array_equal(ops, ref[1]) and np.array_equal(cval, ref[2])
    print('rep %d: failed %d, identical %s, iterations %s' % (rep, int((st != 0).sum()), same, np.bincount(ops - 3)))
    assert same and np.all(st == 0)
print('ok')
