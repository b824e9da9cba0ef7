import os, sys, ctypes as ct
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apm_b200 import _capi, synth
import torch
L = _capi.lib()
L.apm_dev_syrk_bench.argtypes = [ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, ct.POINTER(ct.c_double)]
n, D, B = 768, 8, 256
X, y, th = synth.make_dataset(n, D, seed=0)
eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=B, max_nimp=1)
K = torch.empty(B, n, n, dtype=torch.float64, device='cuda')
eng.kernel_build(synth.bulk_thetas(B, D), out=K)
out = ct.c_double(0)
L.apm_dev_syrk_bench(eng._h, B, 2, 70208, ct.byref(out))
print('syrk %.3f ms' % out.value)
