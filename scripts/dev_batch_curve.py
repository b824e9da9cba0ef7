"""FULL / CACHED estimate time as a function of the number of chains in the call (the sampler's FULL rounds carry
100-200 of a GPU's 256 chains).  LIB=<path> selects another build of the library for A/B runs on the same box."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apm_b200 import _capi, synth
if os.environ.get('LIB'):
    _capi.LIB_PATH = os.path.join(ROOT, os.environ['LIB'])
import torch
n, D, N = int(os.environ.get('N_DATA', 768)), int(os.environ.get('D', 8)), int(os.environ.get('NIMP', 64))
Bs = [int(b) for b in os.environ.get('BS', '32,64,96,128,160,192,224,256').split(',')]
reps = int(os.environ.get('REPS', 10))
Bmax = max(Bs)
X, y, th = synth.make_dataset(n, D, seed=0)
eng = _capi.Engine(X, y, kernel='ard', max_chains=Bmax, n_slots=Bmax, max_nimp=N)
thetas = [synth.bulk_thetas(Bmax, D, seed=s) for s in range(4)]
us = [torch.randn(Bmax, n, N, dtype=torch.float64, device='cuda') for _ in range(2)]
for B in Bs:
    slots = np.arange(B)
    for i in range(3):
        eng.estimate_full(thetas[i % 4][:B], us[i % 2][:B], slots)
    torch.cuda.synchronize(); t = time.perf_counter()
    for i in range(reps):
        out = eng.estimate_full(thetas[i % 4][:B], us[i % 2][:B], slots)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / reps
    torch.cuda.synchronize(); t = time.perf_counter()
    for i in range(reps):
        eng.estimate_cached(slots, us[i % 2][:B])
    torch.cuda.synchronize(); dc = (time.perf_counter() - t) / reps
    print('%-12s B %4d  FULL %7.3f ms %7.0f est/s (%.1f us/chain, iters %.2f)   CACHED %6.3f ms %8.0f est/s' % (
        os.environ.get('TAG', ''), B, dt * 1e3, B / dt, dt * 1e6 / B, (out[1] - 3).mean(), dc * 1e3, B / dc), flush=True)
if os.environ.get('PROFILE_B'):
    B = int(os.environ['PROFILE_B'])
    slots = np.arange(B)
    eng.set_overlap(False)
    eng.profile(True)
    eng.profile_read(reset=True)
    for i in range(reps):
        eng.estimate_full(thetas[i % 4][:B], us[i % 2][:B], slots)
    for name, (ms, cnt) in eng.profile_read(reset=True).items():
        if cnt:
            print('   profile B=%d %-16s %8.3f ms/call %5.1f launches/call' % (B, name, ms / reps, cnt / reps))
    eng.profile(False)
