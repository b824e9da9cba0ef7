import os, sys, ctypes as ct, glob
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib = sys.argv[1]
from apm_b200 import _capi
_capi.LIB_PATH = lib
from apm_b200 import synth
import torch, numpy as np
L = _capi.lib()
L.apm_dev_syrk_bench.argtypes = [ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, ct.POINTER(ct.c_double)]
n, D, B = 768, 8, 256
X, y, th = synth.make_dataset(n, D, seed=0)
eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=B, max_nimp=64)
K = torch.empty(B, n, n, dtype=torch.float64, device='cuda')
thetas = synth.bulk_thetas(B, D)
eng.kernel_build(thetas, out=K)
smem = int(sys.argv[2])
out = ct.c_double(0)
L.apm_dev_syrk_bench(eng._h, B, 5, smem, ct.byref(out))
syrk = B * n**3 / out.value / 1e9
c0 = eng.dev_chol_bench(B, reps=5, mode=0); c1 = eng.dev_chol_bench(B, reps=5, mode=1)
u = torch.randn(B, n, 64, dtype=torch.float64, device='cuda')
import time
eng.estimate_full(thetas, u, np.arange(B))
torch.cuda.synchronize(); t = time.time()
for _ in range(5): eng.estimate_full(thetas, u, np.arange(B))
torch.cuda.synchronize(); dt = (time.time() - t) / 5
print('%-28s syrk %.1f TF/s | chol flow %.3f ms (%.1f TF/s) step %.3f ms (%.1f TF/s) | FULL %.2f ms -> %.0f est/s' % (
    os.path.basename(lib), syrk, c0, B*n**3/3/c0/1e9, c1, B*n**3/3/c1/1e9, dt*1e3, B/dt), flush=True)
