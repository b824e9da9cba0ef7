"""Loader for the UNMODIFIED reference (matt-graham/auxiliary-pm-mcmc) under Python 3.12.

TEST INFRASTRUCTURE ONLY -- never imported by the product package.  Only usable in the build
container, where /root/reference exists; on the GPU box `available()` is False and callers must
fall back to the committed fixtures in tests/golden/ or to oracle/apm_oracle.py.

Zero reference source lines are edited.  Four import shims are installed before import
(SURVEY.md Appendix C):
  1. `scipy.misc.logsumexp`   -> scipy.special.logsumexp    (gpdemo/estimators.py:14)
  2. stub `matplotlib.pyplot` (gpdemo/utils.py:16 imports it at module import time)
  3. `mcmc_updates` alias     -> auxpm.mcmc_updates          (auxpm/samplers.py:11, py2 implicit
                                                              relative import)
  4. `gpdemo.kernels`         -> oracle/_ref/kernels*.so re-cythonized from the unmodified
                                 gpdemo/kernels.pyx by oracle/build_ref.sh (the shipped Cython-0.22
                                 kernels.c does not compile against CPython 3.12).
"""
import glob
import importlib.util
import os
import sys
import types

REFERENCE_DIR = os.environ.get('APM_REFERENCE_DIR', '/root/reference')
_REF_BUILD = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')


def ref_kernels_path():
    so = sorted(glob.glob(os.path.join(_REF_BUILD, 'kernels*.so')))
    return so[0] if so else None


def load_ref_kernels():
    """Import only the compiled reference Cython kernel module (this one travels to the GPU box)."""
    path = ref_kernels_path()
    if path is None:
        return None
    if 'apm_ref_kernels' in sys.modules:
        return sys.modules['apm_ref_kernels']
    spec = importlib.util.spec_from_file_location('kernels', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules['apm_ref_kernels'] = mod
    return mod


def available():
    return (os.path.isfile(os.path.join(REFERENCE_DIR, 'gpdemo', 'estimators.py'))
            and ref_kernels_path() is not None)


_loaded = None


def load_reference():
    """Return a namespace with the reference modules: kernels, lpa, est, utils, mu, smp."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError('reference not available (need %s and oracle/_ref/kernels*.so; '
                           'run oracle/build_ref.sh)' % REFERENCE_DIR)
    import scipy.special
    try:
        import scipy.misc as scipy_misc
    except Exception:  # scipy.misc removed in newer scipy: provide a stub module
        scipy_misc = types.ModuleType('scipy.misc')
        sys.modules['scipy.misc'] = scipy_misc
        import scipy
        scipy.misc = scipy_misc
    scipy_misc.logsumexp = scipy.special.logsumexp                      # shim 1
    if 'matplotlib' not in sys.modules:                                  # shim 2
        m = types.ModuleType('matplotlib')
        m.pyplot = types.ModuleType('matplotlib.pyplot')
        sys.modules['matplotlib'] = m
        sys.modules['matplotlib.pyplot'] = m.pyplot
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import gpdemo
    kernels = load_ref_kernels()
    sys.modules['gpdemo.kernels'] = gpdemo.kernels = kernels             # shim 4
    import auxpm.mcmc_updates as mu
    sys.modules['mcmc_updates'] = mu                                     # shim 3
    import gpdemo.latent_posterior_approximations as lpa
    import gpdemo.estimators as est
    import gpdemo.utils as utils
    import auxpm.samplers as smp
    _loaded = types.SimpleNamespace(kernels=kernels, lpa=lpa, est=est, utils=utils, mu=mu, smp=smp)
    return _loaded
