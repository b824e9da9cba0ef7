import os, sys, ctypes as ct
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apm_b200 import _capi
L = _capi.lib()
L.apm_dev_dmma_sweep.argtypes = [ct.c_int, ct.c_int, ct.c_int, ct.POINTER(ct.c_double)]
for nacc in (1, 4, 16):
    row = []
    for w in (4, 8, 12, 16, 24, 32):
        out = ct.c_double(0)
        L.apm_dev_dmma_sweep(0, w, nacc, ct.byref(out))
        row.append('%dw: %.1f' % (w, out.value))
    print('nacc=%2d  TF/s by warps/SM:  %s' % (nacc, '   '.join(row)))
