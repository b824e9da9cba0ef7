"""Lock-step batched drivers for the auxiliary pseudo-marginal samplers: B independent chains advance
together so that every log-ML estimate they need becomes one batched C-ABI call.

The reference runs chains one after another (notebook cell 14: `for c in range(n_chain)`) and each chain
is scalar Python control flow around `log_f_estimator` (auxpm/samplers.py, auxpm/mcmc_updates.py).  Here
every chain is a generator that *yields* its next estimate request -- ('full', theta) or ('cached', u) --
and is resumed with the value; the scheduler gathers the requests of all live chains, issues one
`apm_estimate_full` and one `apm_estimate_cached` call per round, and hands the results back.  Slice /
elliptical-slice shrink loops therefore stay data dependent per chain (no masking tricks): a chain that
needs three shrink steps simply takes part in three rounds.

Per chain the arithmetic, the order of random draws and the cache hand-over
(current / proposed slot, swapped on acceptance: smp.py:413-417, 1079-1086) are those of the single-chain
classes in `apm_b200.samplers`, so with `rng='parity'` (one `numpy.random.RandomState(seed)` per chain, u
drawn on the host) chain c reproduces the single-chain / reference trace for the same seed, independently of
how many chains share the batch or how they are sharded over GPUs.  `rng='device'` keeps u in HBM
(torch CUDA generator, not stream-identical to numpy) for throughput runs.

`rng='native'` hands the whole run to the native sampler of the C ABI (`apm_sampler_*`, csrc/sampler.cuh): the same chain
logic as a C++ state machine per chain, Philox normals generated on the device straight into the estimator's layout, the
asynchronous FULL / CACHED schedule without the interpreter.  `rng='philox'` runs THIS module's generators on the same
Philox streams (apm_b200.philox, numpy on the host), so the native sampler can be checked draw for draw against the
chain logic that is pinned to the reference.

Chain groups: given a LIST of backends (one engine context each) the chains are split into that many contiguous
groups, each scheduled by its own host thread on its own CUDA stream.  A FULL round of one group (GPU-bound, the GIL
is released inside the C-ABI call) then overlaps the Python scheduling, the torch tensor work and the cheap CACHED
rounds of the others.  Per-chain results are unchanged in parity mode (every chain owns its RandomState).
"""
import numpy as np

from . import utils

TWO_PI = 2. * np.pi


class ChainFailure(Exception):
    """A chain hit one of the reference's exceptions (status code of the C ABI)."""

    def __init__(self, status):
        Exception.__init__(self, 'chain failed with status %d' % status)
        self.status = status


class EngineBackend(object):
    """Batched estimate calls on an apm_b200._capi.Engine.  `u` arguments are lists of per-chain arrays
    (numpy (n, N) on the host in parity mode, torch CUDA tensors in device mode)."""

    def __init__(self, engine):
        self.engine = engine

    def _stack(self, us):
        if isinstance(us[0], np.ndarray):
            return np.ascontiguousarray(np.stack(us))
        import torch
        return torch.stack(us).contiguous()

    def full(self, thetas, us, slots):
        return self.engine.estimate_full(np.asarray(thetas), self._stack(us), slots)

    def cached(self, slots, us):
        return self.engine.estimate_cached(slots, self._stack(us))

    def laplace_lml(self, thetas):
        return self.engine.laplace_lml(np.asarray(thetas))


class _Chain(object):
    """State of one chain + its generator.  `local` = position inside its chain group (slot numbering of the
    group's engine context)."""

    def __init__(self, index, seed, theta_init, local=None, philox_streams=False):
        self.index = index
        local = index if local is None else local
        if philox_streams:
            from .philox import PhiloxStream
            self.prng = PhiloxStream(seed)
        else:
            self.prng = np.random.RandomState(seed)
        self.theta = np.array(theta_init, dtype=np.float64)
        self.u = None
        self.log_f = None
        self.cur_slot = 2 * local
        self.prop_slot = 2 * local + 1
        self.n_reject = [0, 0]
        self.n_cubic_ops = 0
        self.n_full = 0
        self.n_cached = 0
        self.failed = None


class BatchedAPMSampler(object):
    """B chains of one of the composite samplers in lock-step.

    method: 'mi+mh', 'ess+mh', 'mi+rdss', 'ess+rdss' (u-update + theta-update) or 'pmmh' (fresh u inside every
    estimate, smp.py:159-262 with the notebooks' main-phase closure).
    log_prior(theta) -> float is added to every estimate (the notebooks' closure, nb cell 12).
    prop_scales: random-walk scales of the MH theta-update (copied per chain; see SURVEY App. B on the
    reference's aliasing quirk).  slice_width: w of the random-direction slice update (max_steps_out = 0).
    """

    def __init__(self, backend, n_data, n_imp, n_theta, method, log_prior, seeds, prop_scales=None, slice_width=1.,
                 max_slice_iters=1000, rng='parity', device=None, full_batch_frac=0.8, async_full=False,
                 async_batch_frac=0.5):
        if method not in ('mi+mh', 'ess+mh', 'mi+rdss', 'ess+rdss', 'pmmh'):
            raise ValueError('unknown method %r' % method)
        if rng not in ('parity', 'device', 'philox', 'native'):
            raise ValueError("rng must be 'parity', 'philox', 'device' or 'native'")
        self.backends = list(backend) if isinstance(backend, (list, tuple)) else [backend]
        self.backend = self.backends[0]
        self.n, self.N, self.P = int(n_data), int(n_imp), int(n_theta)
        self.method = method
        self.log_prior = log_prior
        self.seeds = list(seeds)
        self.B = len(self.seeds)
        self.prop_scales = None if prop_scales is None else np.array(prop_scales, dtype=np.float64)
        self.slice_width = float(slice_width)
        self.max_slice_iters = int(max_slice_iters)
        self.rng = rng
        self.device = device
        # scheduling policy: FULL estimates are latency-bound for small batches, so they are held back until this
        # fraction of the live chains is waiting for one (or nobody has a cheap CACHED request left)
        self.full_batch_frac = float(full_batch_frac)
        # device-RNG mode only: FULL estimates run on a worker thread (own CUDA stream) while this thread keeps serving
        # the CACHED requests of the other chains through a companion context that shares the engine's cache slots
        self.async_full = bool(async_full)
        self.async_batch_frac = float(async_batch_frac)
        self._async_state = None
        self._gens = [None] * len(self.backends)
        if rng == 'device':
            import torch
            self._torch = torch
            for g in range(len(self.backends)):
                self._gens[g] = torch.Generator(device=device)
                self._gens[g].manual_seed(int(self.seeds[0]) * 7919 + 17 + 104729 * g)
        self._gen = self._gens[0]

    # ---- random draws -----------------------------------------------------------------------------
    def _draw_u(self, ch):
        """u_sampler(): n*N standard normals, row-major (nb cell 12: prng.normal(size=(n, N)))."""
        if self.rng in ('parity', 'philox'):
            return ch.prng.normal(size=(self.n, self.N))
        return self._torch.randn(self.n, self.N, dtype=self._torch.float64, device=self.device, generator=self._gen)

    def _ellipse(self, u, v, phi):
        if self.rng in ('parity', 'philox'):
            return u * np.cos(phi) + v * np.sin(phi)              # mu.py:382
        return u * float(np.cos(phi)) + v * float(np.sin(phi))

    # ---- one chain as a generator of estimate requests -------------------------------------------------
    def _u_update_mi(self, ch):
        u_prop = self._draw_u(ch)                                   # mu.py:284-288
        log_f_prop = yield ('cached', u_prop)
        if ch.prng.uniform() < np.exp(log_f_prop - ch.log_f):       # mu.py:299-303
            ch.u, ch.log_f = u_prop, log_f_prop
        else:
            ch.n_reject[0] += 1

    def _u_update_ess(self, ch):
        v = self._draw_u(ch)                                        # smp.py:786
        log_y = ch.log_f + np.log(ch.prng.uniform())                # mu.py:373
        phi = ch.prng.uniform() * TWO_PI                            # mu.py:375
        lo, hi = phi - TWO_PI, phi
        for _ in range(self.max_slice_iters):
            u_prop = self._ellipse(ch.u, v, phi)
            log_f_prop = yield ('cached', u_prop)
            if log_f_prop > log_y:
                ch.u, ch.log_f = u_prop, log_f_prop
                return
            if phi < 0:
                lo = phi
            elif phi > 0:
                hi = phi
            else:
                return                                              # slice collapsed (mu.py:391-393)
            phi = lo + ch.prng.uniform() * (hi - lo)
        raise ChainFailure(2)

    def _theta_update_mh(self, ch):
        s = self.prop_scales
        theta_prop = ch.theta + s * np.array([ch.prng.normal() for _ in range(self.P)])   # nb cell 12 prop_sampler
        log_f_prop = yield ('full', theta_prop)
        # symmetric Gaussian random walk: forward and backward proposal densities cancel (mu.py:149-152)
        fwd = -0.5 * np.sum(((theta_prop - ch.theta) / s)**2)
        bwd = -0.5 * np.sum(((ch.theta - theta_prop) / s)**2)
        if ch.prng.uniform() < np.exp(log_f_prop + bwd - ch.log_f - fwd):
            ch.theta, ch.log_f = theta_prop, log_f_prop
            ch.cur_slot, ch.prop_slot = ch.prop_slot, ch.cur_slot   # smp.py:413-417
        else:
            ch.n_reject[1] += 1

    def _theta_update_rdss(self, ch):
        d = ch.prng.normal(size=self.P)                             # nb-rdss cell 12 dir_and_w_sampler
        d /= d.dot(d)**0.5
        w = self.slice_width
        log_y = np.log(ch.prng.uniform()) + ch.log_f                # mu.py:481
        lo = 0. - w * ch.prng.uniform()                             # mu.py:483-484
        hi = lo + w
        base = ch.theta.copy()
        for _ in range(self.max_slice_iters):
            x = lo + (hi - lo) * ch.prng.uniform()                  # mu.py:503
            log_f_prop = yield ('full', base + x * d)
            # every evaluation overwrites the current cache (smp.py:1083-1085)
            ch.cur_slot, ch.prop_slot = ch.prop_slot, ch.cur_slot
            if log_f_prop > log_y:
                ch.theta, ch.log_f = base + x * d, log_f_prop
                return
            if x < 0.:
                lo = x
            elif x > 0.:
                hi = x
            else:
                return
        raise ChainFailure(2)

    def _pmmh_update(self, ch):
        s = self.prop_scales
        theta_prop = ch.theta + s * np.array([ch.prng.normal() for _ in range(self.P)])
        ch.u = self._draw_u(ch)                                     # fresh normals inside the estimate closure
        log_f_prop = yield ('full', theta_prop)
        fwd = -0.5 * np.sum(((theta_prop - ch.theta) / s)**2)
        bwd = -0.5 * np.sum(((ch.theta - theta_prop) / s)**2)
        if ch.prng.uniform() < np.exp(log_f_prop + bwd - ch.log_f - fwd):
            ch.theta, ch.log_f = theta_prop, log_f_prop
        else:
            ch.n_reject[1] += 1

    # ---- device-RNG mode: u, v and the proposals live in [B, n, N] tensors; the generators only yield symbolic
    # requests and the scheduler does the tensor work for all chains of a round at once
    def _dev_u_update_mi(self, ch):
        log_f_prop = yield ('cached_new',)
        if ch.prng.uniform() < np.exp(log_f_prop - ch.log_f):
            ch.log_f = log_f_prop
            ch.accept_u = True
        else:
            ch.n_reject[0] += 1

    def _dev_u_update_ess(self, ch):
        log_y = ch.log_f + np.log(ch.prng.uniform())
        phi = ch.prng.uniform() * TWO_PI
        lo, hi = phi - TWO_PI, phi
        new_v = True
        for _ in range(self.max_slice_iters):
            log_f_prop = yield ('cached_ell', phi, new_v)
            new_v = False
            if log_f_prop > log_y:
                ch.log_f = log_f_prop
                ch.accept_u = True
                return
            if phi < 0:
                lo = phi
            elif phi > 0:
                hi = phi
            else:
                return
            phi = lo + ch.prng.uniform() * (hi - lo)
        raise ChainFailure(2)

    def _dev_pmmh_update(self, ch):
        s = self.prop_scales
        theta_prop = ch.theta + s * ch.prng.normal(size=self.P)
        log_f_prop = yield ('full_newu', theta_prop)
        if ch.prng.uniform() < np.exp(log_f_prop - ch.log_f):     # symmetric random walk
            ch.theta, ch.log_f = theta_prop, log_f_prop
        else:
            ch.n_reject[1] += 1

    def _run_chain_device(self, ch, n_sample, trace):
        trace[0] = ch.theta
        ch.log_f = yield ('full_newu', ch.theta)
        if self.method == 'pmmh':
            for s in range(1, n_sample):
                yield from self._dev_pmmh_update(ch)
                trace[s] = ch.theta
            return
        ch.cur_slot, ch.prop_slot = ch.prop_slot, ch.cur_slot
        u_step = self._dev_u_update_mi if self.method.startswith('mi') else self._dev_u_update_ess
        th_step = self._theta_update_mh if self.method.endswith('mh') else self._theta_update_rdss
        for s in range(1, n_sample):
            yield from u_step(ch)
            yield from th_step(ch)
            trace[s] = ch.theta

    def _run_chain(self, ch, n_sample, trace):
        trace[0] = ch.theta
        if self.method == 'pmmh':
            ch.u = self._draw_u(ch)
            ch.log_f = yield ('full', ch.theta)
            for s in range(1, n_sample):
                yield from self._pmmh_update(ch)
                trace[s] = ch.theta
            return
        ch.u = self._draw_u(ch)                                      # smp.py:377 / 546 / 689 / 825
        ch.log_f = yield ('full', ch.theta)
        ch.cur_slot, ch.prop_slot = ch.prop_slot, ch.cur_slot        # the first estimate's cache is current
        u_step = self._u_update_mi if self.method.startswith('mi') else self._u_update_ess
        th_step = self._theta_update_mh if self.method.endswith('mh') else self._theta_update_rdss
        for s in range(1, n_sample):
            yield from u_step(ch)
            yield from th_step(ch)
            trace[s] = ch.theta

    # ---- the lock-step schedulers ----------------------------------------------------------------------
    def _log_prior_many(self, thetas):
        """log prior of several thetas; uses a vectorised callable when the prior provides one."""
        vec = getattr(self.log_prior, 'many', None)
        if vec is not None:
            return vec(np.asarray(thetas))
        return np.array([self.log_prior(t) for t in thetas])

    def _schedule_parity(self, backend, chains, traces, n_sample, gen=None):
        gens = [self._run_chain(ch, n_sample, traces[c]) for c, ch in enumerate(chains)]
        pending = {}
        for c, g in enumerate(gens):
            pending[c] = next(g)
        rounds = 0
        while pending:
            rounds += 1
            full = [c for c, r in pending.items() if r[0] == 'full']
            cached = [c for c, r in pending.items() if r[0] == 'cached']
            results = {}
            if full and cached and len(full) < self.full_batch_frac * len(pending):
                full = []                                        # serve the cheap requests first
            if full:
                thetas = np.stack([pending[c][1] for c in full])
                us = [chains[c].u for c in full]
                slots = [chains[c].prop_slot for c in full]      # written into the proposal slot
                vals, ops, st = backend.full(thetas, us, slots)
                for j, c in enumerate(full):
                    ch = chains[c]
                    ch.n_full += 1
                    if st[j] != 0:
                        results[c] = ChainFailure(int(st[j]))
                    else:
                        ch.n_cubic_ops += int(ops[j])
                        results[c] = float(vals[j]) + self.log_prior(pending[c][1])
            if cached:
                us = [pending[c][1] for c in cached]
                slots = [chains[c].cur_slot for c in cached]
                vals, st = backend.cached(slots, us)
                for j, c in enumerate(cached):
                    ch = chains[c]
                    ch.n_cached += 1
                    if st[j] != 0:
                        results[c] = ChainFailure(int(st[j]))
                    else:
                        results[c] = float(vals[j]) + self.log_prior(ch.theta)
            held = {c: r for c, r in pending.items() if c not in results}
            pending = self._resume(chains, gens, results)
            pending.update(held)
        return rounds

    def _resume(self, chains, gens, results):
        new_pending = {}
        for c, res in results.items():
            try:
                if isinstance(res, ChainFailure):
                    raise res
                new_pending[c] = gens[c].send(res)
            except StopIteration:
                pass
            except ChainFailure as e:           # the notebooks skip a failing chain (nb cell 14)
                chains[c].failed = e.status
        return new_pending

    def _schedule_device(self, backend, chains, traces, n_sample, gen=None):
        torch = self._torch
        B, n, N = len(chains), self.n, self.N
        gen = self._gen if gen is None else gen
        kw = dict(dtype=torch.float64, device=self.device)
        U = torch.zeros(B, n, N, **kw)        # current auxiliary normals of every chain
        V = torch.zeros(B, n, N, **kw)        # ESS auxiliary draw
        Uprop = torch.zeros(B, n, N, **kw)    # proposals of the current round
        for ch in chains:
            ch.accept_u = False
        gens = [self._run_chain_device(ch, n_sample, traces[c]) for c, ch in enumerate(chains)]
        pending = {c: next(g) for c, g in enumerate(gens)}
        rounds = 0

        def idx_t(lst):
            return torch.tensor(lst, dtype=torch.long, device=self.device)

        while pending:
            rounds += 1
            kinds = {}
            for c, r in pending.items():
                kinds.setdefault(r[0], []).append(c)
            results = {}
            # --- FULL estimates (theta changed); 'full_newu' first draws fresh normals for those chains
            full = kinds.get('full_newu', []) + kinds.get('full', [])
            n_cached_req = len(kinds.get('cached_new', [])) + len(kinds.get('cached_ell', []))
            if full and n_cached_req and len(full) < self.full_batch_frac * len(pending):
                full = []                                        # serve the cheap requests first
            if full:
                newu = [c for c in full if pending[c][0] == 'full_newu']
                if newu:
                    U[idx_t(newu)] = torch.randn(len(newu), n, N, generator=gen, **kw)
                thetas = np.stack([pending[c][1] for c in full])
                slots = [chains[c].prop_slot for c in full]
                u_in = U if len(full) == B and full == list(range(B)) else U.index_select(0, idx_t(full))
                vals, ops, st = backend.engine.estimate_full(thetas, u_in.contiguous(), slots)
                lp = self._log_prior_many(thetas)
                for j, c in enumerate(full):
                    ch = chains[c]
                    ch.n_full += 1
                    if st[j] != 0:
                        results[c] = ChainFailure(int(st[j]))
                    else:
                        ch.n_cubic_ops += int(ops[j])
                        results[c] = float(vals[j]) + float(lp[j])
            # --- CACHED estimates (u proposals)
            cached = kinds.get('cached_new', []) + kinds.get('cached_ell', [])
            if cached:
                mi = kinds.get('cached_new', [])
                if mi:
                    Uprop[idx_t(mi)] = torch.randn(len(mi), n, N, generator=gen, **kw)
                ell = kinds.get('cached_ell', [])
                if ell:
                    fresh = [c for c in ell if pending[c][2]]
                    if fresh:
                        V[idx_t(fresh)] = torch.randn(len(fresh), n, N, generator=gen, **kw)
                    it = idx_t(ell)
                    phis = np.array([pending[c][1] for c in ell])
                    cs = torch.tensor(np.cos(phis), **kw)[:, None, None]
                    sn = torch.tensor(np.sin(phis), **kw)[:, None, None]
                    Uprop[it] = U.index_select(0, it) * cs + V.index_select(0, it) * sn       # mu.py:382
                slots = [chains[c].cur_slot for c in cached]
                vals, st = backend.engine.estimate_cached(slots, Uprop.index_select(0, idx_t(cached)).contiguous())
                lp = self._log_prior_many(np.stack([chains[c].theta for c in cached]))
                for j, c in enumerate(cached):
                    ch = chains[c]
                    ch.n_cached += 1
                    results[c] = ChainFailure(int(st[j])) if st[j] != 0 else float(vals[j]) + float(lp[j])
            held = {c: r for c, r in pending.items() if c not in results}
            pending = self._resume(chains, gens, results)
            pending.update(held)
            acc = [c for c in results if chains[c].accept_u]
            if acc:
                ia = idx_t(acc)
                U[ia] = Uprop.index_select(0, ia)
                for c in acc:
                    chains[c].accept_u = False
        return rounds


    def _schedule_device_async(self, backend, chains, traces, n_sample, gen=None):
        """Device-RNG scheduler with asynchronous FULL estimates.  A FULL round is GPU-throughput-bound and takes ~10 ms;
        meanwhile the chains that are in their u-update only need CACHED estimates (O(n^2 N), latency-bound) and Python
        bookkeeping.  The FULL call therefore runs on a worker thread and its own stream (ctypes releases the GIL), and
        this thread keeps scheduling CACHED rounds on a companion context (apm_create_companion: same cache slots,
        own workspaces).  A chain is never in both: the FULL call writes the proposal slots of its chains, the CACHED
        calls read the current slots of the others."""
        import concurrent.futures
        torch = self._torch
        B, n, N = len(chains), self.n, self.N
        gen = self._gen if gen is None else gen
        eng = backend.engine
        if self._async_state is None:
            self._async_state = {}
        st_ = self._async_state.get(id(eng))             # one companion / stream / worker per engine (chain group)
        if st_ is None:
            st_ = dict(engine=eng, comp=eng.companion(), stream=torch.cuda.Stream(device=self.device),
                       pool=concurrent.futures.ThreadPoolExecutor(max_workers=1, thread_name_prefix='apm-full'))
            self._async_state[id(eng)] = st_
        comp, full_stream, pool = st_['comp'], st_['stream'], st_['pool']
        comp.use_torch_stream()                          # CACHED rounds: this thread's current stream
        eng.set_stream(full_stream.cuda_stream)          # FULL rounds: the worker's stream
        kw = dict(dtype=torch.float64, device=self.device)
        U = torch.zeros(B, n, N, **kw)
        V = torch.zeros(B, n, N, **kw)
        Uprop = torch.zeros(B, n, N, **kw)
        for ch in chains:
            ch.accept_u = False
        gens = [self._run_chain_device(ch, n_sample, traces[c]) for c, ch in enumerate(chains)]
        pending = {c: next(g) for c, g in enumerate(gens)}
        inflight = None                                  # (future, chain list, thetas, log priors, u of the call)
        rounds = 0
        import time
        stats = dict(full_calls=0, full_chains=0, cached_calls=0, cached_chains=0, t_submit=0., t_flight=0., t_total=time.perf_counter())
        self.async_stats = stats                         # of the last (group's) run: scheduling diagnostics

        def idx_t(lst):
            return torch.tensor(lst, dtype=torch.long, device=self.device)

        def run_full(thetas, u_in, slots, ready):
            with torch.cuda.device(self.device), torch.cuda.stream(full_stream):
                full_stream.wait_event(ready)            # u_in was gathered on the scheduler's stream
                return eng.estimate_full(thetas, u_in, slots)

        def harvest(fl):
            fut, full, thetas, lp, _keep = fl
            vals, ops, st = fut.result()
            results = {}
            for j, c in enumerate(full):
                ch = chains[c]
                ch.n_full += 1
                if st[j] != 0:
                    results[c] = ChainFailure(int(st[j]))
                else:
                    ch.n_cubic_ops += int(ops[j])
                    results[c] = float(vals[j]) + float(lp[j])
            return results

        def finish_round(results):
            for c in results:
                pending.pop(c, None)
            pending.update(self._resume(chains, gens, results))
            acc = [c for c in results if chains[c].accept_u]
            if acc:
                ia = idx_t(acc)
                U[ia] = Uprop.index_select(0, ia)
                for c in acc:
                    chains[c].accept_u = False

        def is_cached(r):
            return r[0] in ('cached_new', 'cached_ell')

        while pending or inflight is not None:
            rounds += 1
            # --- a finished FULL call is harvested as soon as it is seen (with no CACHED work left: wait for it); its
            # chains are resumed at once so that those which need another FULL estimate join the next call
            if inflight is not None and (inflight[0].done() or not any(is_cached(r) for r in pending.values())):
                res = harvest(inflight)
                stats['t_flight'] += time.perf_counter() - stats['t_submit']
                inflight = None
                finish_round(res)
            # --- submit the next FULL call: enough chains are waiting for one, or nobody has CACHED work left
            if inflight is None:
                full = [c for c, r in pending.items() if not is_cached(r)]
                if full and (len(full) >= self.async_batch_frac * len(pending) or len(full) == len(pending)):
                    newu = [c for c in full if pending[c][0] == 'full_newu']
                    if newu:
                        U[idx_t(newu)] = torch.randn(len(newu), n, N, generator=gen, **kw)
                    thetas = np.stack([pending[c][1] for c in full])
                    slots = [chains[c].prop_slot for c in full]
                    u_in = U.index_select(0, idx_t(full)).contiguous()
                    ready = torch.cuda.Event()
                    ready.record()
                    fut = pool.submit(run_full, thetas, u_in, slots, ready)
                    stats['t_submit'] = time.perf_counter()
                    stats['full_calls'] += 1
                    stats['full_chains'] += len(full)
                    # the log priors are evaluated while the call runs; u_in is kept alive until it is harvested
                    inflight = (fut, full, thetas, self._log_prior_many(thetas), u_in)
                    for c in full:
                        del pending[c]
            # --- CACHED requests of the chains that are not in flight
            cached = [c for c, r in pending.items() if is_cached(r)]
            if cached:
                results = {}
                mi = [c for c in cached if pending[c][0] == 'cached_new']
                if mi:
                    Uprop[idx_t(mi)] = torch.randn(len(mi), n, N, generator=gen, **kw)
                ell = [c for c in cached if pending[c][0] == 'cached_ell']
                if ell:
                    fresh = [c for c in ell if pending[c][2]]
                    if fresh:
                        V[idx_t(fresh)] = torch.randn(len(fresh), n, N, generator=gen, **kw)
                    it = idx_t(ell)
                    phis = np.array([pending[c][1] for c in ell])
                    cs = torch.tensor(np.cos(phis), **kw)[:, None, None]
                    sn = torch.tensor(np.sin(phis), **kw)[:, None, None]
                    Uprop[it] = U.index_select(0, it) * cs + V.index_select(0, it) * sn       # mu.py:382
                slots = [chains[c].cur_slot for c in cached]
                vals, st = comp.estimate_cached(slots, Uprop.index_select(0, idx_t(cached)).contiguous())
                stats['cached_calls'] += 1
                stats['cached_chains'] += len(cached)
                lp = self._log_prior_many(np.stack([chains[c].theta for c in cached]))
                for j, c in enumerate(cached):
                    chains[c].n_cached += 1
                    results[c] = ChainFailure(int(st[j])) if st[j] != 0 else float(vals[j]) + float(lp[j])
                finish_round(results)
        torch.cuda.current_stream(self.device).synchronize()
        stats['t_total'] = time.perf_counter() - stats['t_total']
        eng.use_torch_stream()                           # hand the engine back on the caller's stream
        return rounds

    def _run_groups(self, schedule, chains, traces, n_sample, bounds):
        """One scheduler thread per chain group; in device mode each on its own CUDA stream (the group's engine is
        bound to it, so the torch tensor work and the engine's kernels of a group stay ordered)."""
        import threading
        G = len(bounds) - 1
        rounds, errors = [0] * G, [None] * G
        if self.rng == 'device' and getattr(self, '_group_streams', None) is None:
            self._group_streams = [self._torch.cuda.Stream(device=self.device) for _ in range(len(self.backends))]

        def work(g):
            lo, hi = bounds[g], bounds[g + 1]
            try:
                if self.rng == 'device':
                    torch = self._torch
                    eng = self.backends[g].engine
                    stream = self._group_streams[g]
                    with torch.cuda.device(self.device), torch.cuda.stream(stream):
                        eng.use_torch_stream()        # stays bound to the group's stream (kept alive by the sampler)
                        rounds[g] = schedule(self.backends[g], chains[lo:hi], traces[lo:hi], n_sample, self._gens[g])
                        stream.synchronize()
                else:
                    rounds[g] = schedule(self.backends[g], chains[lo:hi], traces[lo:hi], n_sample, None)
            except BaseException as e:       # re-raised in the caller's thread
                errors[g] = e

        threads = [threading.Thread(target=work, args=(g,), name='apm-group-%d' % g) for g in range(G)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for e in errors:
            if e is not None:
                raise e
        return max(rounds)

    def get_samples(self, theta_init, n_sample, theta_init_sampler=None):
        """theta_init: (B, n_theta), or None with theta_init_sampler(prng) -> theta drawing each chain's start
        from its own stream right after seeding (as the notebooks do, nb cell 14).  Returns dict(thetas
        (B, n_sample, P), n_reject (B, 2), n_cubic_ops (B,), n_full (B,), n_cached (B,), failed (B,) status
        codes, rounds)."""
        B = self.B
        if self.rng == 'native':
            return self._get_samples_native(theta_init, n_sample, theta_init_sampler)
        G = max(1, min(len(self.backends), B))
        bounds = [(g * B) // G for g in range(G + 1)]            # contiguous chain groups, one per backend
        local = {}
        for g in range(G):
            for c in range(bounds[g], bounds[g + 1]):
                local[c] = c - bounds[g]
        px = self.rng == 'philox'
        if px and theta_init is None:
            raise ValueError("rng='philox' needs explicit theta_init (the Philox streams have no Gamma sampler)")
        if theta_init is None:
            chains = [_Chain(c, self.seeds[c], np.zeros(self.P), local[c], px) for c in range(B)]
            for ch in chains:
                ch.theta = np.array(theta_init_sampler(ch.prng), dtype=np.float64)
        else:
            theta_init = np.asarray(theta_init, dtype=np.float64)
            chains = [_Chain(c, self.seeds[c], theta_init[c], local[c], px) for c in range(B)]
        traces = np.full((B, n_sample, self.P), np.nan)
        schedule = self._schedule_parity
        if self.rng == 'device':
            schedule = self._schedule_device_async if self.async_full else self._schedule_device
        if G == 1:
            rounds = schedule(self.backends[0], chains, traces, n_sample, self._gens[0])
        else:
            rounds = self._run_groups(schedule, chains, traces, n_sample, bounds)
        return dict(thetas=traces, n_reject=np.array([ch.n_reject for ch in chains]),
                    n_cubic_ops=np.array([ch.n_cubic_ops for ch in chains]),
                    n_full=np.array([ch.n_full for ch in chains]), n_cached=np.array([ch.n_cached for ch in chains]),
                    failed=np.array([0 if ch.failed is None else ch.failed for ch in chains]), rounds=rounds)


    def _get_samples_native(self, theta_init, n_sample, theta_init_sampler=None):
        """The whole run inside the C ABI (apm_sampler_run).  The log prior must be the notebooks' log-Gamma prior
        (`make_log_prior`: its hyper-parameters are handed to the native code)."""
        from . import _capi
        prior_ab = getattr(self.log_prior, 'prior_ab', None)
        if prior_ab is None:
            raise ValueError("rng='native' needs a log prior made by make_log_prior (log-Gamma hyper-parameters)")
        if theta_init is None:
            raise ValueError("rng='native' needs explicit theta_init (the Philox streams have no Gamma sampler)")
        ns = getattr(self, '_native', None)
        if ns is None:
            ns = _capi.NativeSampler(self.backend.engine, self.method, self.seeds, self.N, prior_ab, self.prop_scales,
                                     self.slice_width, self.max_slice_iters)
            self._native = ns
        thetas, counts = ns.run(np.asarray(theta_init, dtype=np.float64), n_sample)
        self.async_stats = ns.stats()
        return dict(thetas=thetas, n_reject=counts[:, 0:2].copy(), n_cubic_ops=counts[:, 2].copy(), n_full=counts[:, 3].copy(),
                    n_cached=counts[:, 4].copy(), failed=counts[:, 5].copy(), rounds=self.async_stats['rounds'])


def make_log_prior(D, ard):
    """The notebooks' log-Gamma prior (nb cell 8 + cell 12), same tau prior for every ARD length-scale."""
    from . import synth
    p = synth.prior_params(D)

    def log_prior(theta):
        v = utils.log_gamma_log_pdf(theta[0], p['a_sigma'], p['b_sigma'])
        for t in theta[1:]:
            v = v + utils.log_gamma_log_pdf(t, p['a_tau'], p['b_tau'])
        return v

    def many(thetas):            # (B, P) -> (B,), same terms summed in the same order
        v = utils.log_gamma_log_pdf(thetas[:, 0], p['a_sigma'], p['b_sigma'])
        for k in range(1, thetas.shape[1]):
            v = v + utils.log_gamma_log_pdf(thetas[:, k], p['a_tau'], p['b_tau'])
        return v
    log_prior.many = many
    # shape / rate per theta component, for the native sampler
    log_prior.prior_ab = np.array([[p['a_sigma'], p['b_sigma']]] + [[p['a_tau'], p['b_tau']]] * (D if ard else 1))
    return log_prior
