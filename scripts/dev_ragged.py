import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
from apm_b200 import _capi
import apm_oracle as orc
n, D, N, kind = 65, 2, 65, 'iso'
rs = np.random.RandomState(n)
X = rs.normal(size=(n, D)); y = np.where(rs.uniform(size=n) < 0.5, 1., -1.)
thetas = np.r_[0.3, np.full(1, -0.7)][None] + 0.2 * rs.normal(size=(2, 2))
u = rs.normal(size=(2, n, N)); u2 = rs.normal(size=(2, n, N))
for rep in range(3):
    eng = _capi.Engine(X, y, kernel=kind, max_chains=2, max_nimp=N)
    full, ops, st = eng.estimate_full(thetas, u, [1, 0])
    cached, st2 = eng.estimate_cached([1, 0], u2)
    w = eng.cached_weights([1, 0], u2)
    for b in range(2):
        K = np.empty((n, n)); orc.isotropic_squared_exponential_kernel(K, X, thetas[b])
        est = orc.LogMarginalLikelihoodApproxPosteriorISEstimator(X, y, orc.isotropic_squared_exponential_kernel, orc.laplace_approximation)
        ref, cache = est(u[b], thetas[b]); ref2, _ = est(u2[b], None, cache)
        wr = orc.is_log_weights(u2[b], y, *cache)
        print(rep, b, 'cond %.2e full rel %.2e cached rel %.2e w maxabs %.2e (max|w| %.1f) argmax %d' % (
            np.linalg.cond(K), abs(full[b]-ref)/abs(ref), abs(cached[b]-ref2)/abs(ref2), np.max(np.abs(w[b]-wr)), np.max(np.abs(wr)), np.argmax(np.abs(w[b]-wr))))
    eng.close()
