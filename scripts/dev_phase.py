import os, sys, ctypes as ct
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apm_b200 import _capi
_capi.LIB_PATH = os.path.join(ROOT, 'auxiliary-pm-mcmc_b200', 'libapm_timing.so')
from apm_b200 import synth
n, D, B = 768, 8, 256
X, y, th = synth.make_dataset(n, D, seed=0)
thetas = synth.bulk_thetas(B, D)
eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=B, max_nimp=1)
import torch
K = torch.empty(B, n, n, dtype=torch.float64, device='cuda')
eng.kernel_build(thetas, out=K)
L = _capi.lib()
names = ['wait+srcload', 'panel GEMM', 'stage T,Lkk', 'trsm64', 'store+diag syrk', 'store D+publish', 'potrf64', 'store+publish(d)', 'logdet+inverse']
for mode in (0, 1):
    cyc = (ct.c_ulonglong * 16)(); cnt = (ct.c_ulonglong * 16)()
    L.apm_dev_phase_read(cyc, cnt, 1)
    ms = eng.dev_chol_bench(B, reps=3, mode=mode)
    L.apm_dev_phase_read(cyc, cnt, 1)
    tot = sum(cyc[i] for i in range(9))
    print('mode %d: %.3f ms per chol; CTA-cycles by phase (share of total CTA busy time):' % (mode, ms))
    for i, nm in enumerate(names):
        if cnt[i]:
            print('   %-18s avg %8.0f cyc x %7d  = %5.1f%%' % (nm, cyc[i] / cnt[i], cnt[i], 100. * cyc[i] / tot))
