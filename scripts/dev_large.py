"""n = 8192 (BASELINE config 5 shape): does the path run, how long does a FULL / CACHED estimate take."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from apm_b200 import _capi, synth
n, D, N, B = int(os.environ.get('N_DATA', 8192)), 16, 64, int(os.environ.get('B', 4))
t = time.time(); X, y, th = synth.make_dataset(n, D, seed=0); print('data %.1fs' % (time.time() - t), flush=True)
eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=2 * B, max_nimp=N)
thetas = synth.bulk_thetas(B, D, spread=0.1)
u = torch.randn(B, n, N, dtype=torch.float64, device='cuda')
for rep in range(3):
    torch.cuda.synchronize(); t = time.time()
    out, ops, st = eng.estimate_full(thetas, u, np.arange(B))
    dt = time.time() - t
    flop = sum((o - 3) / 3. + 8. / 3. for o in ops) * float(n)**3
    print('FULL  B=%d n=%d: %.1f ms  (%.2f est/s, %.1f TFLOP/s)  ops %s status %s logml %s' % (B, n, dt * 1e3, B / dt, flop / dt / 1e12, ops, st, out), flush=True)
for rep in range(2):
    t = time.time(); out, st = eng.estimate_cached(np.arange(B), u); dt = time.time() - t
    print('CACHED %.1f ms (%.1f est/s)' % (dt * 1e3, B / dt), flush=True)
