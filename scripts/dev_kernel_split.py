"""Per-kernel-family CUDA-event split of the bench workload's FULL estimate (single lane, overlap off) and of a CACHED estimate."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from apm_b200 import _capi, synth
n, D, N, B = 768, 8, 64, int(os.environ.get('B', 256))
steps = int(os.environ.get('STEPS', 6))
X, y, th = synth.make_dataset(n, D, seed=0)
eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=B, max_nimp=N)
thetas = [synth.bulk_thetas(B, D, seed=s) for s in range(4)]
us = [torch.randn(B, n, N, dtype=torch.float64, device='cuda') for _ in range(2)]
slots = np.arange(B)
for i in range(2):
    eng.estimate_full(thetas[i % 4], us[i % 2], slots)
eng.set_overlap(False)
eng.profile(True)
eng.profile_read(reset=True)
for i in range(steps):
    eng.estimate_full(thetas[i % 4], us[i % 2], slots)
prof = eng.profile_read(reset=True)
tot = sum(ms for ms, _ in prof.values())
print('FULL, B=%d: %.2f ms per step (sum of kernel times)' % (B, tot / steps))
for k, (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print('  %-16s %7.3f ms/step  %5.1f launches/step  %5.1f%%' % (k, ms / steps, cnt / steps, 100 * ms / tot))
for i in range(steps):
    eng.estimate_cached(slots, us[i % 2])
prof = eng.profile_read(reset=True)
tot = sum(ms for ms, _ in prof.values())
print('CACHED, B=%d: %.3f ms per step' % (B, tot / steps))
for k, (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    if cnt:
        print('  %-16s %7.3f ms/step  %5.1f launches/step  %5.1f%%' % (k, ms / steps, cnt / steps, 100 * ms / tot))
