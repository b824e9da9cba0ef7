"""apm_b200 -- B200-native pseudo-marginal likelihood engine for GP-probit APM-MCMC.

Drop-in for the hot path of matt-graham/auxiliary-pm-mcmc (gpdemo.kernels,
gpdemo.latent_posterior_approximations, gpdemo.estimators) behind the reference's own Python
callables; all arithmetic runs in hand-written sm_100a CUDA kernels reached through a C ABI
(include/apm_b200.h, csrc/).  There is no CPU fallback: importing `apm_b200._capi` fails loudly
if the CUDA library has not been built (python __graft_entry__.py build).
"""
__version__ = '0.1.0'
