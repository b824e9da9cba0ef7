"""Experiment: two engines with half the chains each, driven from two host threads on two streams, versus one
engine with all chains (overlap of one half's latency-bound kernels with the other half's DMMA kernels)."""
import os, sys, time, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from apm_b200 import _capi, synth
n, D, N = 768, 8, 64
B = int(os.environ.get('TOTAL_CHAINS', 256))
X, y, th = synth.make_dataset(n, D, seed=0)
thetas = synth.bulk_thetas(B, D)
u = torch.randn(B, n, N, dtype=torch.float64, device='cuda')
K = int(os.environ.get('ENGINES', 2))
steps = 10
engs, streams = [], []
for e in range(K):
    eng = _capi.Engine(X, y, kernel='ard', max_chains=B // K, n_slots=2 * (B // K), max_nimp=N)
    st = torch.cuda.Stream()
    eng.set_stream(st.cuda_stream)
    engs.append(eng); streams.append(st)
def work(e):
    sl = slice(e * (B // K), (e + 1) * (B // K))
    for i in range(steps):
        engs[e].estimate_full(thetas[sl], u[sl], np.arange(B // K))
for rep in range(3):
    torch.cuda.synchronize(); t = time.time()
    ths = [threading.Thread(target=work, args=(e,)) for e in range(K)]
    [x.start() for x in ths]; [x.join() for x in ths]
    torch.cuda.synchronize(); dt = time.time() - t
    print('engines=%d grid_div=%s: %.2f ms per 256-chain step -> %.0f est/s' % (K, os.environ.get('APM_FLOW_GRID_DIV', '1'), dt / steps * 1e3, B * steps / dt), flush=True)
