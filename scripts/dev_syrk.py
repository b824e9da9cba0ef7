import os, sys, ctypes as ct
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apm_b200 import _capi, synth
import torch
L = _capi.lib()
L.apm_dev_syrk_bench.argtypes = [ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, ct.POINTER(ct.c_double)]
n, D = 768, 8
X, y, th = synth.make_dataset(n, D, seed=0)
for B in (256, 16):
    eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=B, max_nimp=1)
    K = torch.empty(B, n, n, dtype=torch.float64, device='cuda')
    eng.kernel_build(synth.bulk_thetas(B, D), out=K)
    for smem, occ in ((70208, 3), (110 * 1024, 2), (200 * 1024, 1)):
        out = ct.c_double(0)
        rc = L.apm_dev_syrk_bench(eng._h, B, 5, smem, ct.byref(out))
        print('B=%d occupancy %d CTA/SM: %.3f ms  %.2f TF/s (rc %d)' % (B, occ, out.value, B * n**3 / out.value / 1e9, rc), flush=True)
    eng.close()
