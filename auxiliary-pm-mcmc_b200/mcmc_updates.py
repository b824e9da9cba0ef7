"""Single-chain MCMC transition functions with the call signatures of `auxpm.mcmc_updates`
(auxpm/mcmc_updates.py).  They are the consumers of the hot path: each takes `log_f_func(x) -> float`
(a closure over the log-ML estimator) and a `numpy.random.RandomState`.  The order in which random
numbers are drawn is part of the contract -- it is what makes accept/reject sequences
reproducible against the reference for a fixed seed -- and is noted on every function.

Written from the algorithm descriptions (Metropolis-Hastings; Murray, Adams & MacKay 2010 for
elliptical slice sampling; Neal 2003 for linear slice sampling), host-side scalar control flow only.
"""
import warnings

import numpy as np

TWO_PI = 2. * np.pi


class MaximumIterationsExceededError(Exception):
    """A slice-sampling bracket did not produce an accepted point within max_slice_iters."""


def _mh_decide(prng, log_ratio, x_curr, log_f_curr, x_prop, log_f_prop):
    """Shared accept/reject: one uniform draw, accept iff U < exp(log_ratio)."""
    if prng.uniform() < np.exp(log_ratio):
        return x_prop, log_f_prop, False
    return x_curr, log_f_curr, True


def metropolis_step(x_curr, log_f_curr, log_f_func, prng, prop_sampler, prop_scales):
    """Symmetric-proposal Metropolis update (mu.py:14-76).
    RNG order: prop_sampler(x_curr, prop_scales) draws, [log_f_func], one uniform.
    Returns (x_next, log_f_next, rejected)."""
    x_prop = prop_sampler(x_curr, prop_scales)
    log_f_prop = log_f_func(x_prop)
    return _mh_decide(prng, log_f_prop - log_f_curr, x_curr, log_f_curr, x_prop, log_f_prop)


def met_hastings_step(x_curr, log_f_curr, log_f_func, prng, prop_sampler, prop_params, log_prop_density):
    """Metropolis-Hastings update with a possibly asymmetric proposal (mu.py:79-156).
    log_prop_density(x_to, x_from, prop_params).  RNG order as metropolis_step."""
    x_prop = prop_sampler(x_curr, prop_params)
    log_f_prop = log_f_func(x_prop)
    fwd = log_prop_density(x_prop, x_curr, prop_params)
    bwd = log_prop_density(x_curr, x_prop, prop_params)
    return _mh_decide(prng, log_f_prop + bwd - log_f_curr - fwd, x_curr, log_f_curr, x_prop, log_f_prop)


def metropolis_indepedence_step(x_curr, log_f_curr, log_f_func, prng, prop_sampler, prop_params=None,
                                log_prop_density=None):
    """Metropolis independence update (mu.py:159-303; the reference's spelling of the name is kept).
    With log_prop_density=None the target is taken relative to the proposal, so the ratio is just
    exp(log_f_prop - log_f_curr).  RNG order: prop_sampler() draws, [log_f_func], one uniform."""
    x_prop = prop_sampler(prop_params) if prop_params else prop_sampler()
    log_f_prop = log_f_func(x_prop)
    log_ratio = log_f_prop - log_f_curr
    if log_prop_density:
        args = (prop_params,) if prop_params else ()
        fwd = log_prop_density(x_prop, *args)
        bwd = log_prop_density(x_curr, *args)
        log_ratio = log_f_prop + bwd - log_f_curr - fwd
    return _mh_decide(prng, log_ratio, x_curr, log_f_curr, x_prop, log_f_prop)


def elliptical_slice_step(x_curr, log_f_curr, log_f_func, prng, gaussian_sample, max_slice_iters=1000):
    """Elliptical slice sampling for a target N(x; 0, I) * f(x) (mu.py:311-400).
    RNG order: uniform (slice height), uniform (initial angle), then one uniform per shrink.
    Returns (x_next, log_f_next)."""
    log_y = log_f_curr + np.log(prng.uniform())
    phi = prng.uniform() * TWO_PI
    lo, hi = phi - TWO_PI, phi
    log_f_prop = None
    for _ in range(max_slice_iters):
        x_prop = x_curr * np.cos(phi) + gaussian_sample * np.sin(phi)
        log_f_prop = log_f_func(x_prop)
        if log_f_prop > log_y:
            return x_prop, log_f_prop
        if phi < 0:
            lo = phi
        elif phi > 0:
            hi = phi
        else:
            warnings.warn('Slice collapsed to current value')
            return x_curr, log_f_curr
        phi = lo + prng.uniform() * (hi - lo)
    raise MaximumIterationsExceededError(
        'Exceed maximum slice iterations: i={0}, phi_min={1}, phi_max={2}, log_f_prop={3}, log_f_curr={4}'
        .format(max_slice_iters, lo, hi, log_f_prop, log_f_curr))


def linear_slice_step(x_curr, log_f_curr, log_f_func, slice_width, prng, max_steps_out=0,
                      max_slice_iters=1000):
    """Univariate slice sampling along a line with optional stepping out and shrinkage (mu.py:403-519).
    RNG order: uniform (slice height), uniform (bracket offset), [uniform + stepping-out evaluations if
    max_steps_out > 0], then one uniform per proposal.  Returns (x_next, log_f_next)."""
    log_y = np.log(prng.uniform()) + log_f_curr
    lo = x_curr - slice_width * prng.uniform()
    hi = lo + slice_width
    if max_steps_out > 0:
        n_down = np.round(prng.uniform() * max_steps_out)
        n_up = max_steps_out - n_down
        k = 0
        while k < n_down and log_y < log_f_func(lo):
            lo -= slice_width
            k += 1
        k = 0
        while k < n_up and log_y < log_f_func(hi):
            hi += slice_width
            k += 1
    log_f_prop = None
    for _ in range(max_slice_iters):
        x_prop = lo + (hi - lo) * prng.uniform()
        log_f_prop = log_f_func(x_prop)
        if log_f_prop > log_y:
            return x_prop, log_f_prop
        if x_prop < x_curr:
            lo = x_prop
        elif x_prop > x_curr:
            hi = x_prop
        else:
            warnings.warn('Slice collapsed to current value')
            return x_curr, log_f_curr
    raise MaximumIterationsExceededError(
        'Exceed maximum slice iterations: i={0}, x_min={1}, x_max={2}, log_f_prop={3}, log_f_curr={4}'
        .format(max_slice_iters, lo, hi, log_f_prop, log_f_curr))
