"""Single-chain auxiliary pseudo-marginal samplers with the class names, constructor arguments and
`get_samples` / `adaptive_run` behaviour of `auxpm.samplers` (auxpm/samplers.py), so code written for the
reference can switch estimator AND sampler package together.  (The reference's own sampler classes also run
unchanged on top of the apm_b200 estimators -- see INTEGRATION.md; these mirrors exist because the
reference package is not importable on a machine that only has this repository.)

Every composite sampler is one u-update followed by one theta-update per iteration:

    u-update      'mi'   Metropolis independence proposal u' ~ N(0, I)        (smp.py:382-393, 694-705)
                  'ess'  elliptical slice sampling on u                       (smp.py:508-516, 780-788)
    theta-update  'mh'   (Metropolis-)Hastings random walk                    (smp.py:394-417, 561-584)
                  'seq'  slice sampling one coordinate at a time              (smp.py:900-923)
                  'rd'   slice sampling along a random direction              (smp.py:986-1004, 1071-1089)
                  'ess'  elliptical slice sampling on theta (Gaussian prior)  (smp.py:1147-1165)

They are all driven by `_ApmChain`, which owns the estimator cache hand-over: a u-update re-uses the
current theta's cached factorisations (O(n^2 N) estimates), a theta-update produces a new cache that
replaces the current one only if the move is accepted (MH) or always (slice: the last evaluation is the
accepted point).  RNG consumption order follows SURVEY.md App. B exactly.
"""
import numpy as np

from . import mcmc_updates as mcmc


def _alloc_trace(theta_init, n_sample, column=False):
    if hasattr(theta_init, 'shape'):
        return np.empty((n_sample, theta_init.shape[0]))
    return np.empty((n_sample, 1)) if column else np.empty(n_sample)


class BaseAdaptiveMHSampler(object):
    """Adaptive tuning of random-walk proposal scales (smp.py:14-156)."""

    def __init__(self, prop_scales):
        self.prop_scales = prop_scales

    def get_samples(self, theta_init, n_sample):
        raise NotImplementedError()

    def adaptive_run(self, theta_init, batch_size, n_batch, low_acc_thr, upp_acc_thr, adapt_factor_func,
                     print_details=False, reject_count_index=-1):
        """Run n_batch batches of batch_size samples; after each batch scale `prop_scales` IN PLACE down /
        up by adapt_factor_func(b, n_batch) when the accept rate leaves [low_acc_thr, upp_acc_thr].
        Returns (thetas, prop_scales per batch, accept_rates)."""
        dim = theta_init.shape[0]
        thetas = np.empty((n_batch * batch_size, dim))
        scales = np.empty((n_batch, self.prop_scales.shape[0]))
        rates = np.empty(n_batch)
        for b in range(n_batch):
            seg = slice(b * batch_size, (b + 1) * batch_size)
            thetas[seg], n_rej = self.get_samples(theta_init, batch_size)
            if hasattr(n_rej, '__len__') and reject_count_index:
                n_rej = n_rej[reject_count_index]
            rates[b] = 1. - n_rej * 1. / batch_size
            theta_init = thetas[seg.stop - 1]
            factor = adapt_factor_func(b, n_batch)
            if rates[b] < low_acc_thr:
                self.prop_scales /= factor
            elif rates[b] > upp_acc_thr:
                self.prop_scales *= factor
            scales[b] = self.prop_scales
            if print_details:
                print('Batch {0}: accept rate {1}, adapt factor {2}'.format(b + 1, rates[b], factor))
        return thetas, scales, rates


class PMMHSampler(BaseAdaptiveMHSampler):
    """Pseudo-marginal Metropolis-Hastings: the estimator is a function of theta only and draws its own
    auxiliary variables (smp.py:159-262)."""

    def __init__(self, log_f_estimator, log_prop_density, prop_sampler, prop_scales, prng):
        super(PMMHSampler, self).__init__(prop_scales)
        self.log_f_estimator = log_f_estimator
        self.do_metropolis_update = log_prop_density is None
        if log_prop_density is not None:
            self.log_prop_density = log_prop_density
        self.prop_sampler = prop_sampler
        self.prng = prng

    def get_samples(self, theta_init, n_sample):
        thetas = _alloc_trace(theta_init, n_sample)
        thetas[0] = theta_init
        log_f = self.log_f_estimator(theta_init)
        n_reject = 0
        for s in range(1, n_sample):
            if self.do_metropolis_update:
                thetas[s], log_f, rej = mcmc.metropolis_step(
                    thetas[s - 1], log_f, self.log_f_estimator, self.prng, self.prop_sampler, self.prop_scales)
            else:
                thetas[s], log_f, rej = mcmc.met_hastings_step(
                    thetas[s - 1], log_f, self.log_f_estimator, self.prng, self.prop_sampler, self.prop_scales,
                    self.log_prop_density)
            n_reject += bool(rej)
        return thetas, n_reject


class _ApmChain(object):
    """State and the two half-updates of one auxiliary pseudo-marginal chain."""

    def __init__(self, sampler, theta_init, u_init):
        self.s = sampler
        self.theta = theta_init
        self.u = u_init if u_init is not None else sampler.u_sampler()
        self.log_f, self.cache = sampler.log_f_estimator(self.u, theta_init)
        self.cache_prop = None

    # ---- u | theta : only u changes, so every evaluation re-uses the current cache
    def _log_f_given_cache(self, u):
        return self.s.log_f_estimator(u, self.theta, self.cache)[0]

    def update_u_mi(self):
        self.u, self.log_f, rejected = mcmc.metropolis_indepedence_step(
            self.u, self.log_f, self._log_f_given_cache, self.s.prng, self.s.u_sampler)
        return rejected

    def update_u_ess(self):
        v = self.s.u_sampler()
        self.u, self.log_f = mcmc.elliptical_slice_step(
            self.u, self.log_f, self._log_f_given_cache, self.s.prng, v, self.s.max_slice_iters)

    # ---- theta | u : every evaluation is a FULL estimate producing a fresh cache
    def _log_f_new_theta(self, theta):
        value, self.cache_prop = self.s.log_f_estimator(self.u, theta)
        return value

    def update_theta_mh(self):
        s = self.s
        if s.do_metropolis_update:
            theta, self.log_f, rejected = mcmc.metropolis_step(
                self.theta, self.log_f, self._log_f_new_theta, s.prng, s.prop_sampler, s.prop_scales)
        else:
            theta, self.log_f, rejected = mcmc.met_hastings_step(
                self.theta, self.log_f, self._log_f_new_theta, s.prng, s.prop_sampler, s.prop_scales,
                s.log_prop_density)
        if not rejected:
            self.theta, self.cache = theta, self.cache_prop
        self.cache_prop = None
        return rejected

    def _line_slice(self, point_of, w):
        """Slice sample x along theta(x) = point_of(x), starting from x = 0; the last evaluated point is
        the accepted one, so its cache becomes current."""
        s = self.s
        x_new, self.log_f = mcmc.linear_slice_step(
            0., self.log_f, lambda x: self._log_f_new_theta(point_of(x)), w, s.prng, s.max_steps_out,
            s.max_slice_iters)
        if self.cache_prop is not None:
            self.cache, self.cache_prop = self.cache_prop, None
        return x_new

    def update_theta_rand_dir(self):
        d, w = self.s.slc_dir_and_w_sampler()
        base = self.theta.copy()
        x_new = self._line_slice(lambda x: base + x * d, w)
        self.theta = base + x_new * d

    def update_theta_seq(self):
        theta = self.theta.copy()
        s = self.s
        for j in range(len(theta)):
            # bracket placed around the coordinate's current value; the other coordinates are those
            # already updated in this sweep (smp.py:908-922)
            x_new, self.log_f = mcmc.linear_slice_step(
                theta[j], self.log_f, lambda x, j=j: self._log_f_new_theta(np.r_[theta[:j], x, theta[j + 1:]]),
                s.ws[j], s.prng, s.max_steps_out, s.max_slice_iters)
            if self.cache_prop is not None:
                self.cache, self.cache_prop = self.cache_prop, None
            theta[j] = x_new
        self.theta = theta

    def update_theta_ess(self):
        v = self.s.theta_sampler()
        self.theta, self.log_f = mcmc.elliptical_slice_step(
            self.theta, self.log_f, self._log_f_new_theta, self.s.prng, v, self.s.max_slice_iters)
        if self.cache_prop is not None:
            self.cache, self.cache_prop = self.cache_prop, None


class _ApmMHMixin(object):
    def _setup_mh(self, log_f_estimator, log_prop_density, prop_sampler, prop_scales, u_sampler, prng):
        self.log_f_estimator = log_f_estimator
        self.do_metropolis_update = log_prop_density is None
        if log_prop_density is not None:
            self.log_prop_density = log_prop_density
        self.prop_sampler = prop_sampler
        self.prop_scales = prop_scales
        self.u_sampler = u_sampler
        self.prng = prng


class APMMetIndPlusMHSampler(BaseAdaptiveMHSampler, _ApmMHMixin):
    """MI update of u + (Metropolis-)Hastings update of theta (smp.py:265-418).
    get_samples returns (thetas, (n_reject_u, n_reject_theta))."""

    def __init__(self, log_f_estimator, log_prop_density, prop_sampler, prop_scales, u_sampler, prng):
        super(APMMetIndPlusMHSampler, self).__init__(prop_scales)
        self._setup_mh(log_f_estimator, log_prop_density, prop_sampler, prop_scales, u_sampler, prng)

    def get_samples(self, theta_init, n_sample, u_init=None):
        thetas = _alloc_trace(theta_init, n_sample)
        thetas[0] = theta_init
        chain = _ApmChain(self, thetas[0], u_init)
        rej_u = rej_theta = 0
        for s in range(1, n_sample):
            chain.theta = thetas[s - 1]
            rej_u += bool(chain.update_u_mi())
            rej_theta += bool(chain.update_theta_mh())
            thetas[s] = chain.theta
        return thetas, (rej_u, rej_theta)


class APMEllSSPlusMHSampler(BaseAdaptiveMHSampler, _ApmMHMixin):
    """Elliptical slice update of u + (Metropolis-)Hastings update of theta (smp.py:421-585).
    get_samples returns (thetas, n_reject_theta)."""

    def __init__(self, log_f_estimator, log_prop_density, prop_sampler, prop_scales, u_sampler, prng,
                 max_slice_iters=1000):
        super(APMEllSSPlusMHSampler, self).__init__(prop_scales)
        self._setup_mh(log_f_estimator, log_prop_density, prop_sampler, prop_scales, u_sampler, prng)
        self.max_slice_iters = max_slice_iters

    def get_samples(self, theta_init, n_sample, u_init=None):
        thetas = _alloc_trace(theta_init, n_sample)
        thetas[0] = theta_init
        chain = _ApmChain(self, thetas[0], u_init)
        n_reject = 0
        for s in range(1, n_sample):
            chain.theta = thetas[s - 1]
            chain.update_u_ess()
            n_reject += bool(chain.update_theta_mh())
            thetas[s] = chain.theta
        return thetas, n_reject


class _ApmSliceBase(object):
    """Common constructor of the samplers whose theta-update is a slice move (smp.py:588-841)."""

    _u_update = 'mi'

    def __init__(self, log_f_estimator, u_sampler, prng, max_steps_out=0, max_slice_iters=1000):
        self.log_f_estimator = log_f_estimator
        self.u_sampler = u_sampler
        self.prng = prng
        self.max_steps_out = max_steps_out
        self.max_slice_iters = max_slice_iters

    def _theta_update(self, chain):
        raise NotImplementedError()

    def get_samples(self, theta_init, n_sample, u_init=None):
        """MI-u variants return (thetas, n_reject_u); ESS-u variants return thetas only
        (smp.py:710, 841)."""
        thetas = _alloc_trace(theta_init, n_sample, column=True)
        thetas[0] = theta_init
        chain = _ApmChain(self, thetas[0], u_init)
        n_reject = 0
        for s in range(1, n_sample):
            chain.theta = thetas[s - 1]
            if self._u_update == 'mi':
                n_reject += bool(chain.update_u_mi())
            else:
                chain.update_u_ess()
            self._theta_update(chain)
            thetas[s] = chain.theta
        return (thetas, n_reject) if self._u_update == 'mi' else thetas


class BaseAPMMetIndPlusSliceSampler(_ApmSliceBase):
    _u_update = 'mi'


class BaseAPMEllSSPlusSliceSampler(_ApmSliceBase):
    _u_update = 'ess'


class APMMetIndPlusSeqSliceSampler(BaseAPMMetIndPlusSliceSampler):
    """MI update of u + coordinate-wise slice sampling of theta with widths ws (smp.py:844-923)."""

    def __init__(self, log_f_estimator, u_sampler, prng, ws, max_steps_out=0, max_slice_iters=1000):
        super(APMMetIndPlusSeqSliceSampler, self).__init__(log_f_estimator, u_sampler, prng, max_steps_out,
                                                           max_slice_iters)
        self.ws = ws

    def _theta_update(self, chain):
        chain.update_theta_seq()


class APMMetIndPlusRandDirSliceSampler(BaseAPMMetIndPlusSliceSampler):
    """MI update of u + slice sampling of theta along a random direction (smp.py:926-1004).
    slc_dir_and_w_sampler() -> (direction, width)."""

    def __init__(self, log_f_estimator, u_sampler, prng, slc_dir_and_w_sampler, max_steps_out=0,
                 max_slice_iters=1000):
        super(APMMetIndPlusRandDirSliceSampler, self).__init__(log_f_estimator, u_sampler, prng, max_steps_out,
                                                               max_slice_iters)
        self.slc_dir_and_w_sampler = slc_dir_and_w_sampler

    def _theta_update(self, chain):
        chain.update_theta_rand_dir()


class APMEllSSPlusRandDirSliceSampler(BaseAPMEllSSPlusSliceSampler):
    """Elliptical slice update of u + random-direction slice sampling of theta (smp.py:1007-1089):
    the paper's recommended combination and BASELINE.json configs 2 and 5."""

    def __init__(self, log_f_estimator, u_sampler, prng, slc_dir_and_w_sampler, max_steps_out=0,
                 max_slice_iters=1000):
        super(APMEllSSPlusRandDirSliceSampler, self).__init__(log_f_estimator, u_sampler, prng, max_steps_out,
                                                              max_slice_iters)
        self.slc_dir_and_w_sampler = slc_dir_and_w_sampler

    def _theta_update(self, chain):
        chain.update_theta_rand_dir()


class APMEllSSPlusEllSSSampler(BaseAPMEllSSPlusSliceSampler):
    """Elliptical slice updates of both u and theta (zero-mean Gaussian prior on theta folded into
    theta_sampler) (smp.py:1092-1165)."""

    def __init__(self, log_f_estimator, u_sampler, theta_sampler, prng, max_slice_iters=1000):
        super(APMEllSSPlusEllSSSampler, self).__init__(log_f_estimator, u_sampler, prng, None, max_slice_iters)
        self.theta_sampler = theta_sampler

    def _theta_update(self, chain):
        chain.update_theta_ess()
