#!/usr/bin/env python
"""bench.py -- headline benchmark of the pseudo-marginal likelihood hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA), one process per GPU
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm on host cores

Workload (config.workload): pima-shaped synthetic GP probit (n=768, D=8, ARD kernel, eps=1e-8), Laplace
importance-sampling estimator with N_imp=64, 256 independent chains per GPU.  One "step" = one batched
FULL log-marginal-likelihood estimate (K(theta) -> chol K -> Newton/Laplace -> covariance -> chol C ->
IS tail) for all chains of the rank, with fresh theta and u every step.  metric = FULL estimates / s.

  value   inputs (u) already resident in HBM; timed with CUDA events, max over ranks
  e2e     the same step through the reference-facing C-ABI call with HOST buffers (pinned u, theta):
          H2D of u and theta and D2H of the results inside the timed region
  roofline  dominant kernel family (k_chol = k_chol_flow, the TMA / mbarrier dataflow Cholesky, fp64 DMMA) timed live with CUDA
            events around every launch of the timed region (apm_profile); peak = fp64 DMMA issue peak measured in this run;
            every kernel family carries its own bound / achieved / frac
  configs   the other BASELINE.json configurations (breast E-SS + RD-SS, N_imp sweep, PM-MH, n = 8192), chains sharded over
            the ranks, keyed under "configs" of the same JSON line; the reference arm carries the CPU counterparts of the
            iterations/s figures
  cpu_baseline  the oracle port (numpy/scipy/OpenBLAS + the reference's own Cython kernel module when
            oracle/_ref is present) timed on this box's host cores on a bounded sample
"""
import argparse
import json
import os
import sys
import threading
import time

os.environ.setdefault('OPENBLAS_NUM_THREADS', '1')     # CPU baseline: 1 BLAS thread per process
os.environ.setdefault('OMP_NUM_THREADS', '1')

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(name='pima-shaped synthetic GP probit (n=768, D=8, ARD kernel, eps=1e-8), Laplace importance-sampling estimator, '
                     'N_imp=64: FULL log-ML estimates with fresh theta and u', n=768, D=8, n_imp=64,
                chains_per_gpu=256, kernel='ard', epsilon=1e-8)
METRIC = 'FULL log-ML estimates/sec (GP probit n=768, N_imp=64)'
UNIT = 'estimates/s'


# ---------------------------------------------------------------------------------------------- helpers
def full_flops(n, D, N, iters_total, chains):
    """Algorithmic flops of FULL estimates: per chain (I/3 + 8/3) n^3 + (8I + 1.5D + 1) n^2 + 2 n^2 N."""
    n3, n2 = float(n)**3, float(n)**2
    return (iters_total / 3. + chains * 8. / 3.) * n3 + (8. * iters_total + chains * (1.5 * D + 1)) * n2 + chains * 2. * n2 * N


class ClockSampler(object):
    """Samples SM clock and throttle reasons of one GPU during the timed region (pynvml, 100 ms period)."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4): 'sw_power_cap',
            getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8): 'hw_slowdown',
            getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
            getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
            getattr(nv, 'nvmlClocksEventReasonHwPowerBrakeSlowdown', 0x80): 'hw_power_brake_slowdown',
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(self.samples)}


def make_inputs(n, D, N, B, seed):
    from apm_b200 import synth
    X, y, _ = synth.make_dataset(n, D, seed=0)
    thetas = [synth.bulk_thetas(B, D, seed=seed + 17 * i) for i in range(3)]
    return X, y, thetas


# ---------------------------------------------------------------------------------------------- CPU legs
def _oracle_estimator(X, y, kernel='ard'):
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import apm_oracle as orc
    import ref_loader
    refk = ref_loader.load_ref_kernels()
    name = 'diagonal_squared_exponential_kernel' if kernel == 'ard' else 'isotropic_squared_exponential_kernel'
    if refk is not None:      # the reference's own compiled Cython builder (oracle/_ref), scalar & GIL-bound
        kf = lambda K, X_, th: getattr(refk, name)(K, X_, th, WORKLOAD['epsilon'])  # noqa: E731
        kind = 'port (numpy/scipy restatement oracle/apm_oracle.py; K build by the reference Cython module oracle/_ref)'
    else:
        kf = lambda K, X_, th: getattr(orc, name)(K, X_, th, WORKLOAD['epsilon'])  # noqa: E731
        kind = 'port (numpy/scipy restatement oracle/apm_oracle.py; K build by oracle/kernels_oracle.c)'
    return orc.LogMarginalLikelihoodApproxPosteriorISEstimator(X, y, kf, orc.laplace_approximation), kind


_W = {}


def _cpu_worker_init():
    """Per-process set-up (outside the timed region): data set, thetas, estimator."""
    w = WORKLOAD
    X, y, thetas = make_inputs(w['n'], w['D'], w['n_imp'], 64, 7)
    _W['X'], _W['y'] = X, y
    _W['est'], _ = _oracle_estimator(X, y)
    _W['thetas'] = thetas[0]


def _cpu_worker(args):
    """One host process: `count` FULL estimates on its own thetas/u with 1 BLAS thread."""
    seed, count = args
    w = WORKLOAD
    rs = np.random.RandomState(seed)
    for i in range(count):
        u = rs.normal(size=(w['n'], w['n_imp']))
        _W['est'](u, _W['thetas'][(seed + i) % 64])
    return count


def _cpu_chain_worker(args):
    """One host process: one single-chain sampler run of the reference's algorithm (oracle/apm_oracle_samplers.py) at the
    pima shape, N_imp = 64, ARD kernel -- the CPU counterpart of the GPU arm's lock-step chains.  Returns
    (iterations, seconds, FULL estimates, CACHED estimates)."""
    method, seed, iters = args
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import apm_oracle as orc
    import apm_oracle_samplers as osm
    from apm_b200 import synth
    w = WORKLOAD
    n, D, N = w['n'], w['D'], w['n_imp']
    est = _W['est']
    prior = synth.prior_params(D)
    lg = orc.log_gamma_log_pdf
    counts = [0, 0]

    def log_prior(th):
        return lg(th[0], prior['a_sigma'], prior['b_sigma']) + sum(lg(t, prior['a_tau'], prior['b_tau']) for t in th[1:])

    def log_f_estimator(u, theta=None, cached=None):
        counts[0 if cached is None else 1] += 1
        v, c = est(u, theta, cached)
        return v + log_prior(theta), c

    prng = np.random.RandomState(seed)
    u_sampler = lambda: prng.normal(size=(n, N))  # noqa: E731
    scales = np.full(D + 1, 0.1)
    prop_sampler = lambda th, s: th + s * prng.normal(size=th.shape[0])  # noqa: E731
    log_prop_density = lambda tp, tc, s: -0.5 * np.sum(((tp - tc) / s)**2)  # noqa: E731

    def dir_and_w():
        d = prng.normal(size=D + 1)
        return d / d.dot(d)**0.5, 1.

    theta0 = _W['thetas'][seed % 64]
    import warnings
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        if method == 'pmmh':
            def main(th):
                counts[0] += 1
                return est(prng.normal(size=(n, N)), th)[0] + log_prior(th)
            osm.pmmh_chain(main, log_prop_density, prop_sampler, scales, prng, theta0, iters + 1)
        else:
            osm.apm_chain(method, log_f_estimator, u_sampler, prng, theta0, iters + 1, dir_and_w_sampler=dir_and_w,
                          log_prop_density=log_prop_density, prop_sampler=prop_sampler, prop_scales=scales)
    return iters, time.perf_counter() - t0, counts[0], counts[1]


def cpu_baseline_settings(budget_s=(8., 6.)):
    """cpu_baseline leg of our arm (rank 0): FULL estimates on the host with (a) one BLAS thread and (b) all cores as BLAS
    threads of one process (the third setting, one single-thread process per core, is the --impl reference arm)."""
    w = WORKLOAD
    X, y, thetas = make_inputs(w['n'], w['D'], w['n_imp'], 64, 99)
    est, kind = _oracle_estimator(X, y)
    rs = np.random.RandomState(5)
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass

    def run(budget):
        est(rs.normal(size=(w['n'], w['n_imp'])), thetas[0][0])          # warm-up
        n_done, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < budget:
            est(rs.normal(size=(w['n'], w['n_imp'])), thetas[0][(n_done + 1) % 64])
            n_done += 1
        return n_done, time.perf_counter() - t0

    n1, t1 = run(budget_s[0])
    settings = {'blas_threads_1': {'value': n1 / t1, 'cores': 1, 'sample': '%d FULL estimates in %.1f s' % (n1, t1)}}
    try:
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=cores, user_api='blas'):
            nc, tc = run(budget_s[1])
        settings['blas_threads_nproc'] = {'value': nc / tc, 'cores': cores, 'sample': '%d FULL estimates in %.1f s' % (nc, tc)}
    except Exception as e:      # threadpoolctl missing / BLAS without a thread control: say so instead of guessing
        settings['blas_threads_nproc'] = {'value': None, 'cores': cores, 'sample': 'not measured: %s' % e}
    best = max((k for k in settings if settings[k]['value']), key=lambda k: settings[k]['value'])
    return {'value': settings[best]['value'], 'unit': UNIT, 'cores': settings[best]['cores'], 'kind': 'port',
            'sample': 'best of two single-process settings (%s): %s; n=%d, D=%d, N_imp=%d; %s; host has %d cores; the third SURVEY '
                      'setting (one single-BLAS-thread process per core) is the --impl reference arm'
                      % (best, settings[best]['sample'], w['n'], w['D'], w['n_imp'], kind, cores),
            'settings': settings}


def run_reference_arm(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on all host cores: one single-BLAS-thread
    process per core (the fastest setting found in the survey), each step = `per_worker` FULL estimates per
    process.  After the timed steps the same pool runs one short single-chain sampler run per core (E-SS + RD-SS and
    PM-MH, the reference's loops restated in oracle/apm_oracle_samplers.py) for the CPU iterations/s."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    per_worker = 2
    _, kind = _oracle_estimator(*make_inputs(8, 2, 1, 1, 0)[:2])
    ctx = mp.get_context('fork')
    chain_legs = {}
    with ctx.Pool(cores, initializer=_cpu_worker_init) as pool:
        for wstep in range(args.warmup):
            pool.map(_cpu_worker, [(1000 + wstep * cores + i, 1) for i in range(cores)])
        t0 = time.perf_counter()
        for step in range(args.steps):
            pool.map(_cpu_worker, [(5000 + step * cores + i, per_worker) for i in range(cores)])
        dt = time.perf_counter() - t0
        if not args.no_configs:
            for method, iters in (('ess+rdss', 4), ('pmmh', 6)):
                tw = time.perf_counter()
                res = pool.map(_cpu_chain_worker, [(method, 300 + i, iters) for i in range(cores)])
                tw = time.perf_counter() - tw
                chain_legs[method] = {
                    'value': sum(r[0] for r in res) / tw, 'unit': 'chain-iterations/s', 'chains': cores, 'iterations': iters,
                    'full_estimates_per_iter': sum(r[2] - 1 for r in res) / float(sum(r[0] for r in res)),
                    'cached_estimates_per_iter': sum(r[3] for r in res) / float(sum(r[0] for r in res)),
                    'method': '%s, one chain per host core (%d processes, OPENBLAS_NUM_THREADS=1), reference loops restated in '
                              'oracle/apm_oracle_samplers.py, wall clock of the whole pool' % (method, cores)}
    total = args.steps * cores * per_worker
    value = total / dt
    w = WORKLOAD
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': w['name'], 'n': w['n'], 'D': w['D'], 'n_imp': w['n_imp'], 'kernel': w['kernel'],
                   'execution': 'CPU arm: %d FULL estimates per step (%d per process)' % (cores * per_worker, per_worker)},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d processes x %d FULL estimates per step, OPENBLAS_NUM_THREADS=1 each; %s; /root/reference is '
                                   'absent on the GPU box, so the arm runs the port (timed beside oracle/ref_loader.py in the build '
                                   'container: the port is ~7 %% faster than the unmodified reference, identical values)'
                                   % (cores, per_worker, kind)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    if chain_legs:
        line['apm_iters_per_s'] = chain_legs['ess+rdss']
        line['configs'] = {'pmmh': chain_legs['pmmh']}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------- GPU arm
class _Dist(object):
    """Rank plumbing: barrier + synchronize, MAX / SUM over ranks."""

    def __init__(self, torch, dist, world, dev):
        self.torch, self.dist, self.world, self.dev = torch, dist, world, dev

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, value, op='max'):
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == 'max' else self.dist.ReduceOp.SUM)
        return float(t.item())

    def timed(self, fn, steps):
        """K steps bracketed by barrier + synchronize, CUDA events on the launching stream, max over ranks."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        outs = [fn(i) for i in range(steps)]
        e1.record()
        self.barrier()
        return self.reduce(e0.elapsed_time(e1)), outs

    def wall(self, fn):
        """Host wall clock of fn between barriers, max over ranks (sampler runs: Python scheduler included)."""
        self.barrier()
        t0 = time.perf_counter()
        out = fn()
        self.barrier()
        return self.reduce(time.perf_counter() - t0), out


def _estimator_rates(D_, eng, n, D, N, B, reps, seed, want_cached=True):
    """FULL (and CACHED) estimates/s of one engine with device-resident u; chains of all ranks."""
    torch = D_.torch
    gen = torch.Generator(device=D_.dev)
    gen.manual_seed(seed)
    from apm_b200 import synth
    u = [torch.randn(B, n, N, dtype=torch.float64, device=D_.dev, generator=gen) for _ in range(2)]
    thetas = [synth.bulk_thetas(B, D, seed=seed + 11 * i) for i in range(2)]
    slots = np.arange(B)
    for i in range(2):
        eng.estimate_full(thetas[i], u[i], slots)
    eng.work_count(reset=True)
    ms_full, outs = D_.timed(lambda i: eng.estimate_full(thetas[i % 2], u[i % 2], slots), reps)
    units = sum(eng.work_count(reset=True)) / float(B * reps)
    out = {'full_estimates_per_s': D_.world * B * reps / (ms_full * 1e-3), 'chains': D_.world * B,
           'newton_iters_mean': float(np.mean([(o[1] - 3).mean() for o in outs])),
           'failed_chains': int(D_.reduce(sum((o[2] != 0).sum() for o in outs), 'sum')),
           'executed_n3_over_3_units_per_estimate': units}
    if want_cached:
        ms_c, _ = D_.timed(lambda i: eng.estimate_cached(slots, u[(i + 1) % 2]), reps)
        out['cached_estimates_per_s'] = D_.world * B * reps / (ms_c * 1e-3)
    return out


SAMPLER_NOTE = {
    'native': 'native sampler of the C ABI (apm_sampler_run, csrc/sampler.cuh): C++ chain state machines, Philox normals and the '
              'ellipse generated on the device in the estimator\'s layout, asynchronous FULL / CACHED schedule',
    'device': 'Python generator scheduler (apm_b200.batched), torch device RNG, asynchronous FULL rounds (worker thread + companion '
              'context for the CACHED rounds)'}


def _sampler_rate(D_, eng, n, D, N, B, method, iters, seed, rank, rng='native', th0=None):
    """chain-iterations/s of B lock-step chains per GPU on `eng`: host wall clock of one get_samples call between barriers (max
    over ranks), which includes the scheduler and the drain of the last iterations."""
    from apm_b200 import batched, synth
    drv = batched.BatchedAPMSampler(batched.EngineBackend(eng), n, N, D + 1, method, batched.make_log_prior(D, True),
                                    [seed + rank * B + c for c in range(B)], prop_scales=np.full(D + 1, 0.1), rng=rng,
                                    device=D_.dev, async_full=True)
    if th0 is None:
        th0 = synth.bulk_thetas(B, D, seed=seed)
    drv.get_samples(th0, 3)          # warm-up (allocations, first-use initialisation)
    dt, out = D_.wall(lambda: drv.get_samples(th0, iters + 1))
    if getattr(drv, '_native', None) is not None:
        drv._native.close()
    return {'value': D_.world * B * iters / dt, 'unit': 'chain-iterations/s', 'iterations': iters, 'chains': D_.world * B,
            'method': method, 'full_estimates_per_iter': float(out['n_full'].mean() - 1) / iters,
            'cached_estimates_per_iter': float(out['n_cached'].mean()) / iters,
            'failed_chains': int(D_.reduce((out['failed'] != 0).sum(), 'sum')),
            'scheduler': SAMPLER_NOTE[rng],
            'timing': 'host wall clock of the whole run between barriers, max over ranks, incl. the scheduler and the drain of the '
                      'last iterations'}


def run_configs(D_, rank, quick):
    """BASELINE.json configurations 2-5, chains sharded over the ranks (independent chains: no data-path collective)."""
    from apm_b200 import _capi, synth
    torch, world, dev = D_.torch, D_.world, D_.dev
    out = {}
    # ---- config 2: breast-shaped, E-SS u + RD-SS theta, 256 lock-step chains per GPU
    n, D, N, B = 682, 9, 64, 256
    X, y, _ = synth.make_dataset(n, D, seed=0)
    eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=2 * B, max_nimp=N, device=dev.index)
    eng.use_torch_stream()
    ent = _estimator_rates(D_, eng, n, D, N, B, 4, 3)
    ent['apm'] = _sampler_rate(D_, eng, n, D, N, B, 'ess+rdss', 20 if quick else 40, 2000, rank)
    ent['workload'] = 'breast-shaped synthetic (n=682, D=9, ARD), Laplace IS N_imp=64, E-SS-u + RD-SS-theta, 256 chains per GPU'
    out['breast_ess_rdss'] = ent
    eng.close()
    # ---- config 3: N_imp sweep, 1024 chains sharded over the ranks (Laplace: the reference has no EP; EP = labelled extension)
    n, D = 768, 8
    B = max(1024 // world, 1)
    Ns = (1, 16, 64, 1024) if quick else (1, 4, 16, 64, 256, 1024)
    X, y, _ = synth.make_dataset(n, D, seed=0)
    eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=B, max_nimp=max(Ns), device=dev.index)
    eng.use_torch_stream()
    sweep = {}
    for N in Ns:
        sweep[str(N)] = _estimator_rates(D_, eng, n, D, N, B, 2, 5 + N)
    eng.set_approximation('ep', 1e-6, 100, 1.0)
    ep_sweep = {}
    for N in Ns:
        ep = _estimator_rates(D_, eng, n, D, N, B, 1 if N > 64 else 2, 77 + N, want_cached=False)
        ep_sweep[str(N)] = {'full_estimates_per_s': ep['full_estimates_per_s'], 'ep_iters_mean': ep['newton_iters_mean'],
                            'failed_chains': ep['failed_chains']}
    eng.close()
    out['nimp_sweep'] = {'workload': 'pima-shaped synthetic (n=768, D=8, ARD), IS estimator, N_imp sweep, 1024 chains sharded over '
                                     '%d GPU(s) (%d per GPU)' % (world, B), 'n_imp': sweep,
                         'ep_extension_n_imp': ep_sweep,
                         'ep_extension_n_imp_64': dict(ep_sweep['64'],
                                                       note='EP posterior approximation: not in the reference (SURVEY App. D), checked '
                                                            'against its own restatement in oracle/; "n_imp" above is the reference\'s Laplace '
                                                            'approximation on the same sweep')}
    # ---- config 4: pseudo-marginal MH, 4096 chains on 8 GPUs = 512 per GPU
    n, D, N, B = 768, 8, 64, 512
    eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=2 * B, max_nimp=N, device=dev.index)
    eng.use_torch_stream()
    ent = _sampler_rate(D_, eng, n, D, N, B, 'pmmh', 10 if quick else 20, 3000, rank)
    ent['workload'] = 'pima-shaped synthetic, pseudo-marginal MH (fresh u inside every estimate), 512 chains per GPU'
    out['pmmh'] = ent
    # the headline sampler (E-SS u + RD-SS theta) with 512 instead of 256 chains per GPU: its FULL calls then carry ~270 chains
    # instead of ~140 and run closer to the full-batch rate
    ent = _sampler_rate(D_, eng, n, D, N, B, 'ess+rdss', 20 if quick else 40, 3500, rank)
    ent['workload'] = 'pima-shaped synthetic, E-SS-u + RD-SS-theta, 512 chains per GPU (the headline apm_iters_per_s uses 256)'
    out['pima_ess_rdss_512_chains'] = ent
    eng.close()
    # ---- config 5: n = 8192, D = 16 ARD, E-SS u + RD-SS theta, 8 chains per GPU (3.5 GB of matrices per chain)
    n, D, N, B = 8192, 16, 64, 8
    X, y, _ = synth.make_dataset(n, D, seed=0)
    eng = _capi.Engine(X, y, kernel='ard', max_chains=B, n_slots=2 * B, max_nimp=N, device=dev.index)
    eng.use_torch_stream()
    ent = _estimator_rates(D_, eng, n, D, N, B, 2, 9)
    n3 = float(n)**3
    ent['executed_tflops'] = ent['full_estimates_per_s'] * ent['executed_n3_over_3_units_per_estimate'] / 3. * n3 / 1e12
    ent['survey_tflops'] = ent['full_estimates_per_s'] * (ent['newton_iters_mean'] / 3. + 8. / 3.) * n3 / 1e12
    ent['apm'] = _sampler_rate(D_, eng, n, D, N, B, 'ess+rdss', 2, 4000, rank)
    ent['workload'] = 'large synthetic GP probit (n=8192, D=16, ARD), Laplace IS N_imp=64, E-SS-u + RD-SS-theta, 8 chains per GPU'
    out['large_n8192'] = ent
    eng.close()
    return out


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from apm_b200 import _capi

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device visible -- the product path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    D_ = _Dist(torch, dist, world, dev)
    barrier, timed = D_.barrier, D_.timed

    w = WORKLOAD
    n, D, N, B = w['n'], w['D'], w['n_imp'], args.chains or w['chains_per_gpu']
    X, y, thetas = make_inputs(n, D, N, B, seed=1234 + 1000 * rank)     # every rank: its own chains
    eng = _capi.Engine(X, y, kernel=w['kernel'], epsilon=w['epsilon'], max_chains=B, n_slots=2 * B, max_nimp=N,
                       device=local_rank)
    eng.use_torch_stream()
    slots = [np.arange(B), np.arange(B, 2 * B)]
    gen = torch.Generator(device=dev)
    gen.manual_seed(4321 + rank)
    u_dev = [torch.randn(B, n, N, dtype=torch.float64, device=dev, generator=gen) for _ in range(2)]
    u_host = [torch.empty(B, n, N, dtype=torch.float64).pin_memory() for _ in range(2)]
    for h, d in zip(u_host, u_dev):
        h.copy_(d)
    torch.cuda.synchronize()

    def step_resident(i):
        return eng.estimate_full(thetas[i % 3], u_dev[i % 2], slots[i % 2])

    def step_host(i):
        return eng.estimate_full(thetas[i % 3], u_host[i % 2].numpy(), slots[i % 2])

    def step_cached(i):
        return eng.estimate_cached(slots[i % 2], u_dev[i % 2])

    for i in range(args.warmup):
        step_resident(i)
    for i in range(2):
        step_host(i)

    # ---- timed region (device-resident inputs), clocks sampled
    sampler = ClockSampler(local_rank)
    eng.launch_count(reset=True)
    eng.work_count(reset=True)
    sampler.start()
    ms_total, outs = timed(step_resident, args.steps)
    clocks = sampler.stop()
    launches = eng.launch_count(reset=True)
    work_value = eng.work_count(reset=True)
    bad = int(sum((o[2] != 0).sum() for o in outs))
    iters_total = float(sum((o[1] - 3).sum() for o in outs))          # Newton iterations over all chains & steps
    chains_done = B * args.steps

    # ---- roofline pass: the same K steps again with CUDA events around every launch (apm_profile) and the
    # stream overlap switched off, so that a kernel's event time is its own duration and not a time-share
    eng.set_overlap(False)
    eng.profile(True)
    eng.profile_read(reset=True)
    eng.work_count(reset=True)
    ms_prof, outs_p = timed(step_resident, args.steps)
    prof = eng.profile_read(reset=True)
    chol_units, syrk_units = eng.work_count(reset=True)   # chain-Choleskys / M' builds actually executed (n^3/3 each)
    eng.profile(False)
    eng.set_overlap(True)
    iters_prof = float(sum((o[1] - 3).sum() for o in outs_p))

    # ---- end to end through the host-buffer C-ABI call
    ms_e2e, _ = timed(step_host, args.steps)
    # ---- the O(n^2 N) cached estimate (u-updates), for context
    ms_cached, _ = timed(step_cached, max(args.steps, 3))

    # ---- APM-MCMC iterations/s: ESS-u + RD-SS-theta in lock-step over the same chains (device-resident u)
    apm_iters = args.apm_iters
    apm = _sampler_rate(D_, eng, n, D, N, B, 'ess+rdss', apm_iters, 1000, rank, 'native', thetas[0])
    apm_py = None if args.no_configs else _sampler_rate(D_, eng, n, D, N, B, 'ess+rdss', min(apm_iters, 40), 1000, rank, 'device', thetas[0])

    # ---- diagnostics gather over NCCL (per-chain log-ML of the last step): the only collective of the path
    last = torch.from_numpy(outs[-1][0]).to(dev)
    if world > 1:
        gathered = [torch.empty_like(last) for _ in range(world)]
        dist.all_gather(gathered, last)
        all_logml = torch.cat(gathered).cpu().numpy()
    else:
        all_logml = last.cpu().numpy()

    # ---- one full-batch launch of the plain factorisation, timed alone (chol(K) of the current K matrices): the kernel without
    # the Newton-round extras (fused forward substitution, M' accumulation, V store) that the k_chol family's average includes
    plain_ms = eng.dev_chol_bench(B, reps=5, mode=0) if rank == 0 else None
    fused_fwd = os.environ.get('APM_NO_FUSED_FWD') is None
    eng.close()

    configs = None if args.no_configs else run_configs(D_, rank, args.quick_configs)

    if rank == 0:
        peak_dmma = _capi.measure_fp64_peak(0, local_rank)
        peak_dfma = _capi.measure_fp64_peak(1, local_rank)
        peaks_file = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        hbm_peak, hbm_src = 6650., 'fallback (B200_PROFILING.md)'
        if os.path.isfile(peaks_file):
            hbm_peak, hbm_src = float(json.load(open(peaks_file))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        n3, n2 = float(n)**3, float(n)**2
        tri = 8. * n * (n + 64) / 2.            # bytes of the lower 64-blocks of one n x n fp64 matrix
        flops = {   # algorithmic flops per kernel family over the roofline pass (this rank)
            # chain-Choleskys actually factored: chol(K), chol(B) per B-space Newton round, chol(M') -- the hybrid Newton
            # round makes chol(M') the last iteration's factorisation, so most chains run I + 1, not I + 2 of them
            # ... and the M' = I + L_K^T W L_K builds (n^3/3 each) that k_chol_flow<true> accumulates inside the factorisation
            'k_chol': (chol_units + syrk_units) * n3 / 3.,
            # factored covariance: chol(C) itself is never formed (factored cache), so the TRSM family is
            # only the n^2 N solve of the importance-sampling tail
            'k_trsm_rows': chains_done * n2 * N,
            'k_gemm_tri': chains_done * n2 * N,
        }
        hbm_bytes = {  # algorithmic bytes of the bandwidth-bound families over the roofline pass
            'k_build_K': chains_done * 8. * n2,
            # triangular mat-vec of the M-space Newton rounds: f' = L_K mu~ (k_l_matvec_rev); the other one, t' = P L_K^T b, is
            # accumulated inside k_chol_flow<true, true> unless the fused forward substitution is switched off
            'k_matvec': syrk_units * tri * (1. if fused_fwd else 2.),
            # s = L^-T L^-1 t: the forward substitution runs inside k_chol_flow's diagonal tasks, k_trsv2 reads the factor once
            # (backward substitution) plus the nb explicit 64 x 64 inverse diagonal blocks
            'k_trsv2': iters_prof * (tri + (n + 63) // 64 * 64. * 64. * 8.) * (1. if fused_fwd else 2.),
            # k_is_logw reads F, Zf, U^T (3 x 8 n N) per chain; the log-sum-exp stage is O(N)
            'k_is_epilogue': chains_done * 24. * n * N,
            # u[n][N] -> U^T[Npad][np] (read + write); the anti-transposed factor V of the cache is written by k_chol_flow<true> itself
            'k_transpose_u': chains_done * 16. * n * N,
            # O(n) vector kernels: prep reads f, y and writes W, W^1/2, b, t; finish reads f', f and writes f (9 vectors / iteration)
            'k_newton_vec': iters_prof * 9. * 8. * n,
        }
        kern = {}
        for name, (ms, cnt) in prof.items():
            if cnt == 0:
                continue
            ent = {'ms_total': ms, 'launches': cnt, 'share_of_step': ms / ms_prof}
            if name in flops:
                ent.update(bound='tensor', achieved=flops[name] / (ms * 1e-3) / 1e12, unit='TFLOP/s')
                ent['frac'] = ent['achieved'] / peak_dmma
            elif name in hbm_bytes:
                ent.update(bound='hbm', achieved=hbm_bytes[name] / (ms * 1e-3) / 1e9, unit='GB/s')
                ent['frac'] = ent['achieved'] / hbm_peak
            else:       # 'misc': queue initialisation, masks, copies of O(B) integers -- launch-latency bound
                ent.update(bound='launch', achieved=cnt / (ms * 1e-3), unit='launches/s', frac=None)
            kern[name] = ent
        dom = max((k for k in kern if k in flops), key=lambda k: kern[k]['ms_total'])
        d = kern[dom]
        traffic, traffic_src = None, 'no ncu capture on record for this kernel / shape'
        tfile = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
        if os.path.isfile(tfile):
            try:
                rec = json.load(open(tfile)).get('%s:n=%d:chains=%d' % (dom, n, B))
                if rec:
                    traffic, traffic_src = rec['dram_bytes_per_launch'], rec['source']
            except Exception:
                pass
        roofline = {
            'bound': 'tensor', 'kernel': dom, 'achieved': d['achieved'], 'peak': peak_dmma, 'unit': 'TFLOP/s',
            'frac': d['frac'],
            # dram__bytes_read.sum + dram__bytes_write.sum of ONE full-batch launch of the dominant kernel, from the
            # `ncu --set full` capture named in traffic_source (profiles/ncu_traffic.json; never measured under this run)
            'traffic': traffic, 'traffic_source': traffic_src,
            'algorithmic_bytes_per_launch': B * 2. * tri,
            'achieved_per_launch_gflop': flops[dom] / d['launches'] / 1e9,
            # the same kernel as ONE plain full-batch factorisation (chol K of all chains) timed alone with CUDA events, 5 repetitions
            'plain_full_batch_launch': {'ms': plain_ms, 'achieved': B * n3 / 3. / (plain_ms * 1e-3) / 1e12, 'unit': 'TFLOP/s',
                                        'frac': B * n3 / 3. / (plain_ms * 1e-3) / 1e12 / peak_dmma,
                                        'note': 'the family figure above averages every k_chol_flow launch of the step: Newton rounds with the fused '
                                                'forward substitution and inverse diagonal blocks, the M\' rounds (accumulation from L_K + second, '
                                                'anti-transposed store), straggler rounds of a few chains and empty launches'},
            'avg_launch_ms': d['ms_total'] / d['launches'],
            'measured': 'CUDA events around every launch of a second pass of the same %d steps, stream overlap off '
                        '(%.2f ms/step; the timed `value` pass runs with the chol(K) / M-space overlap on and no per-launch events)'
                        % (args.steps, ms_prof / args.steps),
            'peak_source': 'fp64 DMMA (mma.sync m8n8k4.f64) issue peak measured in this run by apm_measure_fp64_peak; '
                           'MEASURED_PEAKS.json has no fp64 entry (bf16 only). DFMA peak %.1f TFLOP/s. HBM peak %.0f GB/s %s'
                           % (peak_dfma, hbm_peak, hbm_src),
            # SURVEY §8(d) algorithmic work (F_full, which counts TRSM + SYRK + chol(C) = 7/3 n^3 for the covariance)
            # and the work actually executed (the factored covariance needs n^3: 4/3 n^3 less per estimate)
            'whole_step': {'achieved': full_flops(n, D, N, iters_total, chains_done) / (ms_total * 1e-3) / 1e12,
                           'unit': 'TFLOP/s', 'frac': full_flops(n, D, N, iters_total, chains_done) / (ms_total * 1e-3) / 1e12 / peak_dmma,
                           'executed_tflops': (full_flops(n, D, N, iters_total, chains_done) - (iters_total + 8. * chains_done) * n3 / 3.
                                               + (work_value[0] + work_value[1]) * n3 / 3.) / (ms_total * 1e-3) / 1e12,
                           'note': 'achieved = SURVEY F_full / time (the reference algorithm\'s minimal operation count: I chol(B) + TRSM + SYRK + 2 potrf); '
                                   'executed_tflops replaces its (I/3 + 8/3) n^3 by the n^3/3 units actually run (apm_work_count: %.2f chain-Choleskys '
                                   'and %.2f M\' builds per estimate; factored covariance, factored cache, hybrid Newton round)'
                                   % (work_value[0] / chains_done, work_value[1] / chains_done)},
            'kernels': kern,
        }
        cpu = cpu_baseline_settings() if not args.no_cpu_baseline else None
        value = world * chains_done / (ms_total * 1e-3)
        e2e_val = world * chains_done / (ms_e2e * 1e-3)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': w['name'], 'n': n, 'D': D, 'n_imp': N,
                       'kernel': w['kernel'], 'chains_per_gpu': B, 'parallelism': 'independent chains sharded over %d GPU(s), %d per GPU per step' % (world, B),
                       'l2_policy': 'inputs larger than L2: u 100.7 MB/step alternating between two buffers, per-chain matrices 1.2 GB each',
                       'execution': 'one host thread per GPU; per factorisation one persistent TMA / mbarrier dataflow launch over all chains; '
                                    'Newton rounds queued under device-side masks (one host round trip per estimate)',
                       'newton_iters_mean': iters_total / chains_done, 'failed_chains': bad},
            'e2e': {'value': e2e_val, 'unit': UNIT, 'ms_per_step': ms_e2e / args.steps,
                    'h2d_bytes_per_step': int(B * n * N * 8 + B * (D + 1) * 8), 'd2h_bytes_per_step': int(B * (8 + 4 + 4))},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': roofline,
            'cpu_baseline': cpu,
            'cached_estimates_per_s': world * B * max(args.steps, 3) / (ms_cached * 1e-3),
            # E-SS u + RD-SS theta (smp.py:800-841), lock-step chains on the engine of the timed steps
            'apm_iters_per_s': apm,
            'apm_iters_per_s_python_scheduler': apm_py,
            'configs': configs,
            'diagnostics_gather': {'collective': 'nccl all_gather' if world > 1 else 'none (1 GPU)',
                                   'chains': int(all_logml.shape[0]), 'mean_logml': float(np.nanmean(all_logml))},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='apm_b200', choices=['apm_b200', 'reference'])
    ap.add_argument('--chains', type=int, default=0, help='chains per GPU (default 256)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='skip the BASELINE configurations 2-5 (and the CPU sampler legs of the reference arm)')
    ap.add_argument('--quick-configs', action='store_true', help='shorter sampler runs / fewer sweep points in the configurations')
    ap.add_argument('--apm-iters', type=int, default=200, help='iterations of the batched ESS+RDSS sampler leg (chains drain at the end of a run: short runs understate the steady state)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == '__main__':
    sys.exit(main())
