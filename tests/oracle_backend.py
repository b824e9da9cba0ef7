"""Test-only stand-in for apm_b200.batched.EngineBackend that evaluates batched requests with the CPU oracle
(one estimator call per chain).  Lets the lock-step scheduler be checked on a machine without a GPU."""
import numpy as np

import apm_oracle as orc


class OracleBackend(object):
    def __init__(self, X, y, kind='iso', eps=1e-8):
        base = orc.isotropic_squared_exponential_kernel if kind == 'iso' else orc.diagonal_squared_exponential_kernel
        self.est = orc.LogMarginalLikelihoodApproxPosteriorISEstimator(
            X, y, lambda K, X_, th: base(K, X_, th, eps), orc.laplace_approximation)
        self.slots = {}

    def full(self, thetas, us, slots):
        vals, ops, st = [], [], []
        for th, u, s in zip(thetas, us, slots):
            before = self.est.n_cubic_ops
            try:
                v, cache = self.est(np.asarray(u), th)
                self.slots[int(s)] = cache
                vals.append(v); ops.append(self.est.n_cubic_ops - before); st.append(0)
            except orc.MaximumIterationsExceededError:
                vals.append(np.nan); ops.append(0); st.append(2)
            except np.linalg.LinAlgError:
                vals.append(np.nan); ops.append(0); st.append(1)
        return np.array(vals), np.array(ops), np.array(st)

    def cached(self, slots, us):
        vals = [self.est(np.asarray(u), None, self.slots[int(s)])[0] for s, u in zip(slots, us)]
        return np.array(vals), np.zeros(len(vals), dtype=int)
