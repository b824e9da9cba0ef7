"""Host-side checks that need no GPU: the C-ABI library loads, exports every symbol include/apm_b200.h
declares, and the product path fails loudly (no CPU fallback) when no GPU is present."""
import ctypes as ct
import os
import re

import numpy as np
import pytest

from conftest import ROOT, have_gpu


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'apm_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(apm_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_header_symbols():
    from apm_b200 import _capi
    lib = _capi.lib()
    names = header_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), 'libapm_b200.so does not export %s' % name
    assert set(names) == set(_capi.EXPORTS)
    assert b'sm_100a' in lib.apm_version()


def test_no_cpu_fallback_without_gpu():
    if have_gpu():
        pytest.skip('a GPU is present')
    from apm_b200 import _capi, estimators, kernels, latent_posterior_approximations as lpa
    X = np.random.RandomState(0).normal(size=(10, 2))
    y = np.sign(np.random.RandomState(1).normal(size=10))
    with pytest.raises(_capi.ApmError, match='no GPU'):
        _capi.Engine(X, y)
    est = estimators.LogMarginalLikelihoodApproxPosteriorISEstimator(
        X, y, kernels.isotropic_squared_exponential_kernel, lpa.laplace_approximation)
    with pytest.raises(_capi.ApmError, match='no GPU'):
        est(np.zeros((10, 1)), np.zeros(2))
    with pytest.raises(_capi.ApmError, match='no GPU'):
        kernels.isotropic_squared_exponential_kernel(np.empty((10, 10)), X, np.zeros(2))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'auxiliary-pm-mcmc_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'apm_oracle' not in text and 'ref_loader' not in text, f
                assert '/root/reference' not in text, f


def test_estimator_argument_errors():
    from apm_b200 import estimators, kernels, latent_posterior_approximations as lpa
    X = np.zeros((5, 2))
    y = np.ones(5)
    est = estimators.LogMarginalLikelihoodApproxPosteriorISEstimator(
        X, y, kernels.isotropic_squared_exponential_kernel, lpa.laplace_approximation)
    with pytest.raises(ValueError):
        est(np.zeros((5, 1)))           # neither theta nor cached_results (estimators.py:201-202)
    pm = estimators.LogMarginalLikelihoodPriorMCEstimator(X, y, kernels.isotropic_squared_exponential_kernel)
    with pytest.raises(ValueError):
        pm(np.zeros((5, 1)))
    est.n_cubic_ops = 5
    est.reset_cubic_op_count()
    assert est.n_cubic_ops == 0
