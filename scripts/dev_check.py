"""Developer check on a GPU box: stage-by-stage comparison of the CUDA path with the oracle,
plus first timings.  Not part of the test-suite (tests/ holds the real parity tests)."""
import os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
from apm_b200 import _capi, synth
import apm_oracle as orc
import scipy.linalg as la

def rel(a, b):
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))

def stage(name, fn):
    try:
        t = time.time(); fn(); print('[ok] %s (%.2fs)' % (name, time.time() - t), flush=True)
    except Exception:
        print('[FAIL] %s' % name); traceback.print_exc(); sys.stdout.flush()

def check_shape(n, D, N, B, kernel):
    print('=== n=%d D=%d N=%d B=%d kernel=%s' % (n, D, N, B, kernel), flush=True)
    X, y, theta_true = synth.make_dataset(n, D, seed=3)
    rs = np.random.RandomState(11)
    P = D + 1 if kernel == 'ard' else 2
    base = theta_true[:P]
    thetas = base[None] + 0.2 * rs.normal(size=(B, P))
    u = rs.normal(size=(B, n, N))
    eng = _capi.Engine(X, y, kernel=kernel, max_chains=B, max_nimp=N)
    kf = orc.diagonal_squared_exponential_kernel if kernel == 'ard' else orc.isotropic_squared_exponential_kernel
    Kref = np.empty((B, n, n))
    for b in range(B): kf(Kref[b], X, thetas[b])
    def s_kernel():
        K = eng.kernel_build(thetas)
        print('   K maxrel', rel(K, Kref), 'bit-identical frac', float(np.mean(K == Kref)))
    stage('kernel_build', s_kernel)
    def s_laplace():
        f, C, lml, ops, st = eng.laplace(Kref, calc_cov=True, calc_lml=True)
        for b in range(B):
            fr, Cr, lr, opr = orc.laplace_approximation(Kref[b], y, calc_cov=True, calc_lml=True)
            print('   chain %d: f rel %.2e  C rel %.2e  lml %.12g vs %.12g  ops %d/%d st %d' % (
                b, rel(f[b], fr), rel(C[b], Cr), lml[b], lr, ops[b], opr, st[b]))
    stage('laplace', s_laplace)
    def s_full():
        full, ops, st = eng.estimate_full(thetas, u, np.arange(B))
        u2 = rs.normal(size=(B, n, N))
        cached, st2 = eng.estimate_cached(np.arange(B), u2)
        for b in range(B):
            est = orc.LogMarginalLikelihoodApproxPosteriorISEstimator(X, y, kf, orc.laplace_approximation)
            ref, cache = est(u[b], thetas[b])
            ref2, _ = est(u2[b], None, cache)
            print('   chain %d: full %.12f ref %.12f rel %.2e | cached rel %.2e | ops %d/%d st %d' % (
                b, full[b], ref, abs(full[b]-ref)/abs(ref), abs(cached[b]-ref2)/abs(ref2), ops[b], est.n_cubic_ops, st[b]))
            if b == 0:
                Kc, Cc, fp, ld = eng.slot_export(0)
                print('   slot0: K_chol rel %.2e C_chol rel %.2e f rel %.2e logdets %s vs %s' % (
                    rel(Kc, cache[0]), rel(Cc, cache[1]), rel(fp, cache[2]), ld,
                    (np.log(cache[0].diagonal()).sum(), np.log(cache[1].diagonal()).sum())))
    stage('estimate_full/cached', s_full)
    def s_lml():
        lml, ops, st = eng.laplace_lml(thetas)
        for b in range(min(B, 2)):
            e = orc.LogMarginalLikelihoodLaplaceEstimator(X, y, kf)
            r = e(thetas[b]); print('   lml %.12f ref %.12f ops %d/%d' % (lml[b], r, ops[b], e.n_cubic_ops))
    stage('laplace_lml', s_lml)
    eng.close()

def timing(n, D, N, B, reps=3):
    import torch
    print('=== timing n=%d D=%d N=%d B=%d' % (n, D, N, B), flush=True)
    X, y, theta_true = synth.make_dataset(n, D, seed=0)
    thetas = synth.bulk_thetas(B, D)
    eng = _capi.Engine(X, y, kernel='ard', max_chains=B, max_nimp=N)
    u = torch.randn(B, n, N, dtype=torch.float64, device='cuda')
    torch.cuda.synchronize()
    for r in range(reps):
        t = time.time(); out, ops, st = eng.estimate_full(thetas, u, np.arange(B)); dt = time.time() - t
        print('   FULL  %.2f ms -> %.1f est/s  (ops mean %.2f, bad %d, launches %d)' % (dt*1e3, B/dt, ops.mean(), (st != 0).sum(), eng.launch_count(True)), flush=True)
    for r in range(reps):
        t = time.time(); out, st = eng.estimate_cached(np.arange(B), u); dt = time.time() - t
        print('   CACHED %.2f ms -> %.1f est/s' % (dt*1e3, B/dt), flush=True)
    eng.close()

if __name__ == '__main__':
    print(_capi.lib().apm_version().decode())
    stage('peak', lambda: print('   DMMA peak %.2f TF/s, DFMA peak %.2f TF/s' % (_capi.measure_fp64_peak(0), _capi.measure_fp64_peak(1))))
    check_shape(100, 4, 8, 2, 'ard')
    check_shape(100, 4, 8, 2, 'iso')
    check_shape(200, 5, 70, 3, 'ard')
    check_shape(768, 8, 64, 2, 'ard')
    timing(768, 8, 64, 32)
    timing(768, 8, 64, 256)
