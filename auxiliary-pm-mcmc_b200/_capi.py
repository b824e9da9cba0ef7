"""ctypes binding of the C ABI in include/apm_b200.h (libapm_b200.so, built by
`python __graft_entry__.py build`).  There is NO CPU fallback: a missing library or a missing GPU
raises immediately."""
import ctypes as ct
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libapm_b200.so')

# every symbol include/apm_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    'apm_version', 'apm_last_error', 'apm_create', 'apm_destroy', 'apm_set_stream', 'apm_synchronize',
    'apm_set_overlap', 'apm_set_newton', 'apm_set_approximation', 'apm_ep', 'apm_get_info', 'apm_kernel_build', 'apm_kernel_grad', 'apm_laplace', 'apm_estimate_full',
    'apm_estimate_cached', 'apm_estimate_cached_weights', 'apm_laplace_lml', 'apm_estimate_prior_mc',
    'apm_slot_export', 'apm_slot_import', 'apm_slot_factor', 'apm_slot_copy', 'apm_profile', 'apm_profile_read', 'apm_launch_count', 'apm_work_count', 'apm_create_companion', 'apm_dev_chol_bench', 'apm_measure_fp64_peak',
    'apm_sampler_create', 'apm_sampler_run', 'apm_sampler_stats', 'apm_sampler_destroy',
]

KERNEL_ISO, KERNEL_ARD = 0, 1
CHAIN_OK, CHAIN_CHOL_K, CHAIN_NEWTON_MAXIT, CHAIN_CHOL_C, CHAIN_NONFINITE, CHAIN_CHOL_B = range(6)

_lib = None


class ApmError(RuntimeError):
    pass


def lib():
    """Load the CUDA library (once).  Fails loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            'apm_b200: %s not found -- the CUDA extension has not been built '
            '(run `python __graft_entry__.py build`).  There is no CPU fallback.' % LIB_PATH)
    L = ct.CDLL(LIB_PATH)
    dp, ip, vp = ct.POINTER(ct.c_double), ct.POINTER(ct.c_int), ct.c_void_p
    L.apm_version.restype = ct.c_char_p
    L.apm_last_error.restype = ct.c_char_p
    L.apm_create.argtypes = [vp, vp, ct.c_int, ct.c_int, ct.c_int, ct.c_double, ct.c_int, ct.c_int, ct.c_int,
                             ct.c_int, ct.POINTER(vp)]
    L.apm_destroy.argtypes = [vp]
    L.apm_create_companion.argtypes = [vp, ct.c_int, ct.c_int, ct.POINTER(vp)]
    L.apm_set_stream.argtypes = [vp, ct.c_uint64]
    L.apm_synchronize.argtypes = [vp]
    L.apm_set_overlap.argtypes = [vp, ct.c_int]
    L.apm_set_newton.argtypes = [vp, ct.c_double, ct.c_int]
    L.apm_set_approximation.argtypes = [vp, ct.c_int, ct.c_double, ct.c_int, ct.c_double]
    L.apm_ep.argtypes = [vp, vp, ct.c_int, ct.c_int, ct.c_int, vp, vp, ct.c_int, vp, vp, vp, vp]
    L.apm_get_info.argtypes = [vp] + [ip] * 7
    L.apm_kernel_build.argtypes = [vp, vp, ct.c_int, ct.c_int, ct.c_double, vp, ct.c_int]
    L.apm_kernel_grad.argtypes = [vp, vp, ct.c_int, ct.c_int, vp, ct.c_int]
    L.apm_laplace.argtypes = [vp, vp, ct.c_int, ct.c_int, ct.c_int, ct.c_int, vp, vp, ct.c_int, vp, vp, vp]
    L.apm_estimate_full.argtypes = [vp, vp, vp, ct.c_int, ct.c_int, ct.c_int, vp, vp, vp, vp]
    L.apm_estimate_cached.argtypes = [vp, vp, vp, ct.c_int, ct.c_int, ct.c_int, vp, vp]
    L.apm_estimate_cached_weights.argtypes = [vp, vp, vp, ct.c_int, ct.c_int, ct.c_int, vp]
    L.apm_laplace_lml.argtypes = [vp, vp, ct.c_int, vp, vp, vp]
    L.apm_estimate_prior_mc.argtypes = [vp, vp, vp, vp, ct.c_int, ct.c_int, ct.c_int, vp, vp]
    L.apm_slot_export.argtypes = [vp, ct.c_int, vp, vp, vp, vp]
    L.apm_slot_import.argtypes = [vp, ct.c_int, vp, vp, vp]
    L.apm_slot_factor.argtypes = [vp, ct.c_int, vp, vp, ct.c_int, vp, vp]
    L.apm_slot_copy.argtypes = [vp, vp, vp, ct.c_int]
    L.apm_profile.argtypes = [vp, ct.c_int]
    L.apm_profile_read.argtypes = [vp, ct.c_int, ct.c_char_p, vp, vp, ct.c_int]
    L.apm_launch_count.argtypes = [vp, ct.c_int]
    L.apm_launch_count.restype = ct.c_int64
    L.apm_work_count.argtypes = [vp, ct.POINTER(ct.c_int64), ct.c_int]
    L.apm_dev_chol_bench.argtypes = [vp, ct.c_int, ct.c_int, ct.c_int, dp]
    L.apm_measure_fp64_peak.argtypes = [ct.c_int, ct.c_int, dp]
    L.apm_sampler_create.argtypes = [vp, ct.c_int, ct.c_int, ct.c_int, vp, vp, vp, ct.c_double, ct.c_int, ct.POINTER(vp)]
    L.apm_sampler_run.argtypes = [vp, vp, ct.c_int, vp, vp]
    L.apm_sampler_stats.argtypes = [vp, vp]
    L.apm_sampler_destroy.argtypes = [vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is ct.c_int and name not in ('apm_version', 'apm_last_error', 'apm_launch_count'):
            fn.restype = ct.c_int
    _lib = L
    return L


def check(rc):
    if rc != 0:
        msg = lib().apm_last_error().decode()
        kinds = {1: 'invalid argument', 2: 'CUDA error', 3: 'out of device memory', 4: 'no GPU'}
        raise ApmError('apm_b200: %s: %s' % (kinds.get(rc, 'error %d' % rc), msg))


def _ptr(a):
    """void* of a numpy array (host) -- caller keeps the array alive."""
    return a.ctypes.data_as(ct.c_void_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def measure_fp64_peak(kind=0, device=0):
    out = ct.c_double(0.)
    check(lib().apm_measure_fp64_peak(device, kind, ct.byref(out)))
    return out.value


class Engine(object):
    """Thin object wrapper of one apm_ctx.  Bulk device inputs are passed as torch CUDA tensors
    (torch is used only for buffer handling); host inputs as numpy arrays."""

    def __init__(self, X, y, kernel='ard', epsilon=1e-8, max_chains=1, n_slots=None, max_nimp=1, device=0):
        L = lib()
        X = f64(X)
        y = f64(y)
        if X.ndim != 2 or y.shape != (X.shape[0],):
            raise ValueError('X must be (n, D) and y (n,)')
        self.n, self.D = X.shape
        self.kind = {'iso': KERNEL_ISO, 'ard': KERNEL_ARD}[kernel] if isinstance(kernel, str) else int(kernel)
        self.n_theta = self.D + 1 if self.kind == KERNEL_ARD else 2
        self.max_chains = int(max_chains)
        self.n_slots = int(n_slots) if n_slots is not None else 2 * self.max_chains
        self.max_nimp = int(max_nimp)
        self.device = int(device)
        self.epsilon = float(epsilon)
        h = ct.c_void_p()
        check(L.apm_create(_ptr(X), _ptr(y), self.n, self.D, self.kind, self.epsilon, self.max_chains,
                           self.n_slots, self.max_nimp, self.device, ct.byref(h)))
        self._h = h
        self._L = L

    def companion(self, max_chains=None, max_nimp=None):
        """A context of its own (streams, workspaces) that shares this engine's cache slots: cached estimates on it
        may run while another thread is inside estimate_full on this engine (different slots).  See
        include/apm_b200.h:apm_create_companion."""
        comp = object.__new__(Engine)
        comp.__dict__.update({k: v for k, v in self.__dict__.items() if k not in ('_h', '_companions')})
        comp.max_chains = int(max_chains) if max_chains is not None else self.max_chains
        comp.max_nimp = int(max_nimp) if max_nimp is not None else self.max_nimp
        h = ct.c_void_p()
        check(self._L.apm_create_companion(self._h, comp.max_chains, comp.max_nimp, ct.byref(h)))
        comp._h = h
        comp._parent = self                      # keeps the slot owner alive
        self.__dict__.setdefault('_companions', []).append(comp)
        return comp

    def close(self):
        for comp in self.__dict__.pop('_companions', []):
            comp.close()
        if getattr(self, '_h', None):
            self._L.apm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing
    def set_stream(self, cuda_stream):
        check(self._L.apm_set_stream(self._h, int(cuda_stream)))

    def use_torch_stream(self):
        import torch
        self.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def synchronize(self):
        check(self._L.apm_synchronize(self._h))

    def set_overlap(self, enable=True):
        check(self._L.apm_set_overlap(self._h, 1 if enable else 0))

    def set_newton(self, diff_f_tol=1e-4, max_iters=1000):
        check(self._L.apm_set_newton(self._h, float(diff_f_tol), int(max_iters)))

    def set_approximation(self, kind='laplace', ep_tol=1e-6, ep_max_iters=100, ep_damping=1.0):
        """Posterior approximation of estimate_full: 'laplace' (reference) or 'ep' (extension)."""
        k = {'laplace': 0, 'ep': 1}[kind]
        check(self._L.apm_set_approximation(self._h, k, float(ep_tol), int(ep_max_iters), float(ep_damping)))

    def profile(self, enable=True):
        check(self._L.apm_profile(self._h, 1 if enable else 0))

    def profile_read(self, reset=True):
        """{kernel family: (total ms, launches)} accumulated while profiling was enabled."""
        m = 32
        names = ct.create_string_buffer(32 * m)
        ms = np.zeros(m)
        cnt = np.zeros(m, dtype=np.int64)
        k = self._L.apm_profile_read(self._h, m, names, _ptr(ms), _ptr(cnt), 1 if reset else 0)
        if k < 0:
            raise ApmError('apm_profile_read failed')
        out = {}
        for i in range(k):
            name = names.raw[32 * i:32 * (i + 1)].split(b'\0')[0].decode()
            out[name] = (float(ms[i]), int(cnt[i]))
        return out

    def dev_chol_bench(self, B, reps=5, mode=0):
        out = ct.c_double(0.)
        check(self._L.apm_dev_chol_bench(self._h, int(B), int(reps), int(mode), ct.byref(out)))
        return out.value

    def launch_count(self, reset=False):
        return int(self._L.apm_launch_count(self._h, 1 if reset else 0))

    def work_count(self, reset=False):
        """(chain-Choleskys, M' builds) executed since the last reset, each n^3/3 flops."""
        out = (ct.c_int64 * 2)()
        check(self._L.apm_work_count(self._h, out, 1 if reset else 0))
        return int(out[0]), int(out[1])

    @staticmethod
    def _bulk(a):
        """(pointer, on_device, keepalive) for a numpy array or a torch tensor."""
        if isinstance(a, np.ndarray):
            a = f64(a)
            return _ptr(a), 0, a
        import torch
        if isinstance(a, torch.Tensor):
            if a.dtype != torch.float64 or not a.is_contiguous():
                raise ValueError('device buffers must be contiguous float64 tensors')
            return ct.c_void_p(a.data_ptr()), (1 if a.is_cuda else 0), a
        a = f64(a)
        return _ptr(a), 0, a

    def _theta(self, theta, n_theta=None):
        th = f64(theta)
        if th.ndim == 1:
            th = th[None]
        P = self.n_theta if n_theta is None else n_theta
        if th.shape[1] != P:
            raise ValueError('theta must have %d components per chain' % P)
        return th

    # -- gpdemo.kernels
    def kernel_build(self, theta, kind=None, epsilon=None, out=None):
        k = self.kind if kind is None else kind
        th = self._theta(theta, self.D + 1 if k == KERNEL_ARD else 2)
        B = th.shape[0]
        if out is None:
            out = np.empty((B, self.n, self.n))
        p, dev, keep = self._bulk(out)
        check(self._L.apm_kernel_build(self._h, _ptr(th), B, -1 if kind is None else int(kind),
                                       -1. if epsilon is None else float(epsilon), p, dev))
        return out

    def kernel_grad(self, theta, kind=None, out=None):
        """dK/dtheta_p, (B, n_theta, n, n) -- extension, see include/apm_b200.h:apm_kernel_grad."""
        k = self.kind if kind is None else kind
        P = self.D + 1 if k == KERNEL_ARD else 2
        th = self._theta(theta, P)
        B = th.shape[0]
        if out is None:
            out = np.empty((B, P, self.n, self.n))
        p, dev, keep = self._bulk(out)
        check(self._L.apm_kernel_grad(self._h, _ptr(th), B, -1 if kind is None else int(kind), p, dev))
        return out

    # -- gpdemo.latent_posterior_approximations
    def laplace(self, K, calc_cov=True, calc_lml=False, C_out=None):
        p, dev, keep = self._bulk(K)
        B = 1 if keep.ndim == 2 else keep.shape[0]
        f = np.empty((B, self.n))
        lml = np.empty(B)
        ops = np.empty(B, dtype=np.int32)
        st = np.empty(B, dtype=np.int32)
        C = None
        cp, cdev = None, 0
        if calc_cov:
            C = np.empty((B, self.n, self.n)) if C_out is None else C_out
            cp, cdev, _ = self._bulk(C)
        check(self._L.apm_laplace(self._h, p, dev, B, int(calc_cov), int(calc_lml), _ptr(f), cp, cdev, _ptr(lml),
                                  _ptr(ops), _ptr(st)))
        return f, C, lml, ops, st

    def ep(self, K, calc_cov=True, C_out=None):
        """EP posterior approximation (extension): (f, C, nu, tau, ops, status) for K (B, n, n)."""
        p, dev, keep = self._bulk(K)
        B = 1 if keep.ndim == 2 else keep.shape[0]
        f, nu, tau = np.empty((B, self.n)), np.empty((B, self.n)), np.empty((B, self.n))
        ops = np.empty(B, dtype=np.int32)
        st = np.empty(B, dtype=np.int32)
        C, cp, cdev = None, None, 0
        if calc_cov:
            C = np.empty((B, self.n, self.n)) if C_out is None else C_out
            cp, cdev, _ = self._bulk(C)
        check(self._L.apm_ep(self._h, p, dev, B, int(calc_cov), _ptr(f), cp, cdev, _ptr(nu), _ptr(tau), _ptr(ops), _ptr(st)))
        return f, C, nu, tau, ops, st

    # -- gpdemo.estimators
    def _check_batch(self, keep, sl, B, unique=True):
        """u must be (B, n, N) (or (n, N) for B = 1), slots one per chain, in range and distinct: the C side indexes
        host / device memory with these without further checks."""
        shape = tuple(keep.shape)
        if len(shape) == 2:
            shape = (1,) + shape
        if len(shape) != 3 or shape[0] != B or shape[1] != self.n:
            raise ValueError('u must have shape (%d, %d, N), got %r' % (B, self.n, tuple(keep.shape)))
        N = shape[2]
        if N < 1 or N > self.max_nimp:
            raise ValueError('N = %d importance samples out of range for this engine (max_nimp = %d)' % (N, self.max_nimp))
        if B < 1 or B > self.max_chains:
            raise ValueError('batch of %d chains out of range for this engine (max_chains = %d)' % (B, self.max_chains))
        if sl.shape != (B,):
            raise ValueError('need one slot per chain (%d), got %r' % (B, sl.shape))
        if sl.min() < 0 or sl.max() >= self.n_slots:
            raise ValueError('slot index out of range (n_slots = %d)' % self.n_slots)
        if unique and np.unique(sl).shape[0] != B:
            raise ValueError('slots of one batch must be distinct')
        return N

    def estimate_full(self, theta, u, slots):
        th = self._theta(theta)
        B = th.shape[0]
        p, dev, keep = self._bulk(u)
        sl = i32(np.atleast_1d(slots))
        N = self._check_batch(keep, sl, B)
        out = np.empty(B)
        ops = np.empty(B, dtype=np.int32)
        st = np.empty(B, dtype=np.int32)
        check(self._L.apm_estimate_full(self._h, _ptr(th), p, dev, N, B, _ptr(sl), _ptr(out), _ptr(ops), _ptr(st)))
        return out, ops, st

    def estimate_cached(self, slots, u):
        sl = i32(np.atleast_1d(slots))
        B = sl.shape[0]
        p, dev, keep = self._bulk(u)
        N = self._check_batch(keep, sl, B, unique=False)
        out = np.empty(B)
        st = np.empty(B, dtype=np.int32)
        check(self._L.apm_estimate_cached(self._h, _ptr(sl), p, dev, N, B, _ptr(out), _ptr(st)))
        return out, st

    def cached_weights(self, slots, u):
        sl = i32(np.atleast_1d(slots))
        B = sl.shape[0]
        p, dev, keep = self._bulk(u)
        N = self._check_batch(keep, sl, B, unique=False)
        out = np.empty((B, N))
        check(self._L.apm_estimate_cached_weights(self._h, _ptr(sl), p, dev, N, B, _ptr(out)))
        return out

    def laplace_lml(self, theta):
        th = self._theta(theta)
        B = th.shape[0]
        out = np.empty(B)
        ops = np.empty(B, dtype=np.int32)
        st = np.empty(B, dtype=np.int32)
        check(self._L.apm_laplace_lml(self._h, _ptr(th), B, _ptr(out), _ptr(ops), _ptr(st)))
        return out, ops, st

    def estimate_prior_mc(self, theta, slots, u):
        sl = i32(np.atleast_1d(slots))
        B = sl.shape[0]
        thp = None
        if theta is not None:
            th = self._theta(theta)
            thp = _ptr(th)
        p, dev, keep = self._bulk(u)
        N = self._check_batch(keep, sl, B, unique=theta is not None)
        out = np.empty(B)
        st = np.empty(B, dtype=np.int32)
        check(self._L.apm_estimate_prior_mc(self._h, thp, _ptr(sl), p, dev, N, B, _ptr(out), _ptr(st)))
        return out, st

    # -- slots
    def slot_export(self, slot, want_K=True, want_C=True, want_f=None):
        want_f = want_C if want_f is None else want_f        # a prior-MC cache (chol K only) has neither C_chol nor f_post
        Kc = np.empty((self.n, self.n)) if want_K else None
        Cc = np.empty((self.n, self.n)) if want_C else None
        f = np.empty(self.n) if want_f else None
        ld = np.empty(2)
        check(self._L.apm_slot_export(self._h, int(slot), _ptr(Kc) if want_K else None,
                                      _ptr(Cc) if want_C else None, _ptr(f) if want_f else None, _ptr(ld)))
        return Kc, Cc, f, ld

    def slot_import(self, slot, K_chol, C_chol=None, f_post=None):
        Kc = f64(K_chol)
        Cc = f64(C_chol) if C_chol is not None else None
        fp = f64(f_post) if f_post is not None else None
        check(self._L.apm_slot_import(self._h, int(slot), _ptr(Kc), _ptr(Cc) if Cc is not None else None,
                                      _ptr(fp) if fp is not None else None))

    def slot_factor(self, slot, K, C, f_post):
        kp, kdev, k1 = self._bulk(K)
        cp, cdev, k2 = self._bulk(C)
        if kdev != cdev:
            raise ValueError('K and C must both be host or both be device buffers')
        fp = f64(f_post)
        st = np.zeros(1, dtype=np.int32)
        check(self._L.apm_slot_factor(self._h, int(slot), kp, cp, kdev, _ptr(fp), _ptr(st)))
        return int(st[0])

    def slot_copy(self, src, dst):
        s, d = i32(np.atleast_1d(src)), i32(np.atleast_1d(dst))
        check(self._L.apm_slot_copy(self._h, _ptr(s), _ptr(d), s.shape[0]))


METHODS = {'mi+mh': 0, 'ess+mh': 1, 'mi+rdss': 2, 'ess+rdss': 3, 'pmmh': 4}


class NativeSampler(object):
    """apm_sampler of include/apm_b200.h: B chains of a composite sampler advanced natively on `engine` (device RNG,
    host C++ chain state machines).  prior_ab: (n_theta, 2) shape / rate of the log-Gamma prior of every theta component."""

    def __init__(self, engine, method, seeds, n_imp, prior_ab, prop_scales=None, slice_width=1., max_slice_iters=1000):
        self._L = lib()
        self.engine = engine
        self.B = len(seeds)
        self.P = engine.n_theta
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        prior_ab = f64(prior_ab)
        if prior_ab.shape != (self.P, 2):
            raise ValueError('prior_ab must be (n_theta, 2)')
        ps = None if prop_scales is None else f64(np.broadcast_to(np.asarray(prop_scales, dtype=np.float64), (self.P,)))
        h = ct.c_void_p()
        check(self._L.apm_sampler_create(engine._h, METHODS[method], self.B, int(n_imp), _ptr(seeds), _ptr(prior_ab),
                                         None if ps is None else _ptr(ps), float(slice_width), int(max_slice_iters),
                                         ct.byref(h)))
        self._h = h

    def run(self, theta_init, n_sample):
        theta_init = f64(theta_init)
        if theta_init.shape != (self.B, self.P):
            raise ValueError('theta_init must be (n_chains, n_theta)')
        thetas = np.empty((self.B, int(n_sample), self.P))
        counts = np.zeros((self.B, 6), dtype=np.int64)
        check(self._L.apm_sampler_run(self._h, _ptr(theta_init), int(n_sample), _ptr(thetas), _ptr(counts)))
        return thetas, counts

    def stats(self):
        out = np.zeros(8)
        check(self._L.apm_sampler_stats(self._h, _ptr(out)))
        return dict(full_calls=int(out[0]), full_chains=int(out[1]), cached_calls=int(out[2]), cached_chains=int(out[3]),
                    t_flight=float(out[4]), t_total=float(out[5]), rounds=int(out[6]), t_busy=float(out[7]))

    def close(self):
        if getattr(self, '_h', None):
            self._L.apm_sampler_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
