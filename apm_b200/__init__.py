"""Importable alias for the product package, which lives in the directory `auxiliary-pm-mcmc_b200/`
(a name Python's import statement cannot spell).  `import apm_b200.estimators` resolves to
`auxiliary-pm-mcmc_b200/estimators.py` through the package search path set here."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      'auxiliary-pm-mcmc_b200')
__path__ = [_real]
with open(_os.path.join(_real, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))
del _f
