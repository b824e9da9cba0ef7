"""The notebooks' sampler wiring (experiment_notebooks/*.ipynb cells 8-12) parameterised by which
implementation supplies the estimator and the sampler classes; mirrors oracle/gen_golden.py:build_sampler."""
import numpy as np

from apm_b200 import synth


def build_sampler(method, X, y, N, prng, est_cls, lap_func, iso_kernel, log_gamma_log_pdf, smp, eps=1e-8):
    D = X.shape[1]
    prior = synth.prior_params(D)
    kf = lambda K, X_, th: iso_kernel(K, X_, th, eps)  # noqa: E731
    ml = est_cls(X, y, kf, lap_func)

    def log_prior(theta):
        return (log_gamma_log_pdf(theta[0], prior['a_sigma'], prior['b_sigma']) +
                log_gamma_log_pdf(theta[1], prior['a_tau'], prior['b_tau']))

    def log_f_estimator(u, theta=None, cached_res=None):
        v, c = ml(u, theta, cached_res)
        return v + log_prior(theta), c

    u_sampler = lambda: prng.normal(size=(y.shape[0], N))  # noqa: E731
    prop_sampler = lambda th, s: np.r_[th[0] + s[0] * prng.normal(), th[1] + s[1] * prng.normal()]  # noqa: E731
    log_prop_density = lambda tp, tc, s: -0.5 * (((tp[0] - tc[0]) / s[0])**2 + ((tp[1] - tc[1]) / s[1])**2)  # noqa: E731
    scales = np.array([0.5, 0.5])

    def dir_and_w():
        d = prng.normal(size=2)
        d /= d.dot(d)**0.5
        return d, 1.

    if method == 'mi+mh':
        s = smp.APMMetIndPlusMHSampler(log_f_estimator, log_prop_density, prop_sampler, scales, u_sampler, prng)
    elif method == 'ess+mh':
        s = smp.APMEllSSPlusMHSampler(log_f_estimator, log_prop_density, prop_sampler, scales, u_sampler, prng)
    elif method == 'mi+rdss':
        s = smp.APMMetIndPlusRandDirSliceSampler(log_f_estimator, u_sampler, prng, dir_and_w, 0)
    elif method == 'ess+rdss':
        s = smp.APMEllSSPlusRandDirSliceSampler(log_f_estimator, u_sampler, prng, dir_and_w, 0)
    elif method == 'pmmh':
        main = lambda th: ml(prng.normal(size=(y.shape[0], N)), th)[0] + log_prior(th)  # noqa: E731
        s = smp.PMMHSampler(main, log_prop_density, prop_sampler, scales, prng)
    else:
        raise ValueError(method)
    return s, ml


def run_golden_case(method, N, g, n_iter, **impl):
    X, y = g['X'], g['y']
    prng = np.random.RandomState()
    smp, ml = build_sampler(method, X, y, N, prng, **impl)
    prng.seed(1000 + N)
    theta_init = synth.draw_theta_prior(prng, X.shape[1], ard=False)
    res = smp.get_samples(theta_init, n_iter)
    thetas = res[0] if isinstance(res, tuple) else res
    n_rej = np.atleast_1d(res[1]) if isinstance(res, tuple) else np.zeros(0)
    return thetas, n_rej, ml.n_cubic_ops


def first_divergence(a, b, rtol=1e-7):
    bad = np.where(np.any(np.abs(a - b) > rtol * (1. + np.abs(b)), axis=-1))[0]
    return int(bad[0]) if bad.size else None
