import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apm_b200 import _capi, synth
if os.environ.get('LIB'):
    _capi.LIB_PATH = os.path.join(ROOT, 'auxiliary-pm-mcmc_b200', os.environ['LIB'])
n, D, B = 768, 8, int(os.environ.get('B', 256))
X, y, th = synth.make_dataset(n, D, seed=0)
thetas = synth.bulk_thetas(B, D)
eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=B, max_nimp=1)
import torch
K = torch.empty(B, n, n, dtype=torch.float64, device='cuda')
eng.kernel_build(thetas, out=K)
flop = B * n**3 / 3.
for mode in [int(v) for v in os.environ.get('MODES', '0,16,32').split(',')]:   # chol(K) / chol(B)+inverse blocks / 1 chain in 8 active
    ms = eng.dev_chol_bench(B, reps=5, mode=mode)
    act = 1. / 8 if mode >> 4 == 2 else 1.
    print('%s mode=%d  %.3f ms  %.2f TF/s' % (os.environ.get('TAG', ''), mode, ms, act * flop / ms / 1e9), flush=True)
