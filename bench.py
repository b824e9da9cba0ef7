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
  roofline  dominant kernel family (k_chol = k_chol_dataflow / k_chol_step, fp64 DMMA) timed live with CUDA events around every launch
            of the timed region (apm_profile); peak = fp64 DMMA issue peak measured in this run
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

WORKLOAD = dict(name='pima-shaped synthetic GP probit, Laplace-IS FULL estimate', n=768, D=8, n_imp=64,
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
def _oracle_estimator(X, y):
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import apm_oracle as orc
    import ref_loader
    refk = ref_loader.load_ref_kernels()
    if refk is not None:      # the reference's own compiled Cython builder (oracle/_ref), scalar & GIL-bound
        kf = lambda K, X_, th: refk.diagonal_squared_exponential_kernel(K, X_, th, WORKLOAD['epsilon'])  # noqa: E731
        kind = 'port (numpy/scipy restatement; K build by the reference Cython module oracle/_ref)'
    else:
        kf = lambda K, X_, th: orc.diagonal_squared_exponential_kernel(K, X_, th, WORKLOAD['epsilon'])  # noqa: E731
        kind = 'port (numpy/scipy restatement; K build by oracle/kernels_oracle.c)'
    return orc.LogMarginalLikelihoodApproxPosteriorISEstimator(X, y, kf, orc.laplace_approximation), kind


_W = {}


def _cpu_worker_init():
    """Per-process set-up (outside the timed region): data set, thetas, estimator."""
    w = WORKLOAD
    X, y, thetas = make_inputs(w['n'], w['D'], w['n_imp'], 64, 7)
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


def cpu_baseline_single(budget_s=12.):
    """cpu_baseline leg of our arm: 1 host thread, FULL estimates for about budget_s seconds."""
    w = WORKLOAD
    X, y, thetas = make_inputs(w['n'], w['D'], w['n_imp'], 64, 99)
    est, kind = _oracle_estimator(X, y)
    rs = np.random.RandomState(5)
    est(rs.normal(size=(w['n'], w['n_imp'])), thetas[0][0])          # warm-up
    n_done, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        est(rs.normal(size=(w['n'], w['n_imp'])), thetas[0][(n_done + 1) % 64])
        n_done += 1
    dt = time.perf_counter() - t0
    return {'value': n_done / dt, 'unit': UNIT, 'cores': 1, 'kind': 'port',
            'sample': '%d FULL estimates (n=%d, D=%d, N_imp=%d) in %.1f s, OPENBLAS_NUM_THREADS=1; %s; host has %d cores'
                      % (n_done, w['n'], w['D'], w['n_imp'], dt, kind, os.cpu_count())}


def run_reference_arm(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on all host cores: one single-BLAS-thread
    process per core (the fastest setting found in the survey), each step = `per_worker` FULL estimates per
    process."""
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
    with ctx.Pool(cores, initializer=_cpu_worker_init) as pool:
        for wstep in range(args.warmup):
            pool.map(_cpu_worker, [(1000 + wstep * cores + i, 1) for i in range(cores)])
        t0 = time.perf_counter()
        for step in range(args.steps):
            pool.map(_cpu_worker, [(5000 + step * cores + i, per_worker) for i in range(cores)])
        dt = time.perf_counter() - t0
    total = args.steps * cores * per_worker
    value = total / dt
    w = WORKLOAD
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': '%s; CPU arm: %d FULL estimates per step' % (w['name'], cores * per_worker),
                   'n': w['n'], 'D': w['D'], 'n_imp': w['n_imp'], 'kernel': w['kernel']},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d processes x %d FULL estimates per step, OPENBLAS_NUM_THREADS=1 each; %s'
                                   % (cores, per_worker, kind)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------- GPU arm
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize, CUDA events on the launching stream, max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        outs = [fn(i) for i in range(steps)]
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), outs

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
    from apm_b200 import batched
    drv = batched.BatchedAPMSampler(batched.EngineBackend(eng), n, N, D + 1, 'ess+rdss', batched.make_log_prior(D, True),
                                    [1000 + rank * B + c for c in range(B)], rng='device', device=dev, async_full=True)
    apm_iters = args.apm_iters
    drv.get_samples(thetas[0], 3)          # warm-up (allocations, first-use initialisation)
    barrier()
    t_apm = time.perf_counter()
    apm_out = drv.get_samples(thetas[0], apm_iters + 1)
    barrier()
    t_apm = time.perf_counter() - t_apm
    t_apm_t = torch.tensor([t_apm], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_apm_t, op=dist.ReduceOp.MAX)
    t_apm = float(t_apm_t.item())

    # ---- diagnostics gather over NCCL (per-chain log-ML of the last step): the only collective of the path
    last = torch.from_numpy(outs[-1][0]).to(dev)
    if world > 1:
        gathered = [torch.empty_like(last) for _ in range(world)]
        dist.all_gather(gathered, last)
        all_logml = torch.cat(gathered).cpu().numpy()
    else:
        all_logml = last.cpu().numpy()

    if rank == 0:
        peak_dmma = _capi.measure_fp64_peak(0, local_rank)
        peak_dfma = _capi.measure_fp64_peak(1, local_rank)
        peaks_file = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        hbm_peak, hbm_src = 6650., 'fallback (B200_PROFILING.md)'
        if os.path.isfile(peaks_file):
            hbm_peak, hbm_src = float(json.load(open(peaks_file))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        n3 = float(n)**3
        flops = {   # algorithmic flops per kernel family over the timed region (this rank)
            # chain-Choleskys actually factored: chol(K), chol(B) per B-space Newton round, chol(M') -- the hybrid Newton
            # round makes chol(M') the last iteration's factorisation, so most chains run I + 1, not I + 2 of them
            'k_chol': chol_units * n3 / 3.,
            # factored covariance: M' = I + Y'Y'^T is n^3/3; chol(C) itself is never formed (factored cache),
            # so the TRSM family is only the n^2 N solve of the importance-sampling tail
            'k_trsm_rows': chains_done * float(n)**2 * N,
            'k_syrk_sub': syrk_units * n3 / 3.,                        # M' = I + Y'Y'^T builds executed
            'k_gemm_tri': chains_done * float(n)**2 * N,
        }
        hbm_bytes = {  # algorithmic bytes of the bandwidth-bound families
            'k_build_K': chains_done * 8. * n * n,
            # triangular mat-vecs of the M-space Newton rounds (L_K^T b and L_K mu~, lower tiles of L_K each); the B-space
            # rounds are mat-vec-free (k_fnew_from_s reads O(n) vectors)
            'k_matvec': syrk_units * 2. * 8. * (n * (n + 64) / 2.),
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
            kern[name] = ent
        dom = max((k for k in kern if k in flops), key=lambda k: kern[k]['ms_total'])
        d = kern[dom]
        roofline = {
            'bound': 'tensor', 'kernel': dom, 'achieved': d['achieved'], 'peak': peak_dmma, 'unit': 'TFLOP/s',
            'frac': d['frac'],
            # dram__bytes_read.sum + dram__bytes_write.sum of ONE k_chol_dataflow launch (256 chains, n = 768) from the
            # `ncu --set full` capture summarised in profiles/r1h_ncu_top_kernels_full.md: 4.21 GB + 1.20 GB
            'traffic': 5.41e9 if (dom == 'k_chol' and n == 768 and B == 256) else None,
            'traffic_note': 'bytes per launch from ncu (profiles/r1h_ncu_top_kernels_full.md); minimum (read K, write L) '
                            'is 1.28 GB, the blocked left-looking operand traffic with a working set > L2 is 4.7 GB',
            'achieved_per_launch_gflop': flops[dom] / d['launches'] / 1e9,
            'avg_launch_ms': d['ms_total'] / d['launches'],
            'measured': 'CUDA events around every launch of a second pass of the same %d steps, single lane, stream overlap off '
                        '(%.2f ms/step; the timed `value` pass runs with lanes and overlap on and no per-launch events)' % (args.steps, ms_prof / args.steps),
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
        cpu = cpu_baseline_single() if not args.no_cpu_baseline else None
        value = world * chains_done / (ms_total * 1e-3)
        e2e_val = world * chains_done / (ms_e2e * 1e-3)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': '%s; %d chains per GPU per step' % (w['name'], B), 'n': n, 'D': D, 'n_imp': N,
                       'kernel': w['kernel'], 'chains_per_gpu': B, 'parallelism': 'independent chains sharded over %d GPU(s)' % world,
                       'l2_policy': 'inputs larger than L2: u 100.7 MB/step alternating between two buffers, per-chain matrices 1.2 GB each',
                       'execution': 'apm_estimate_full splits the batch into lanes (up to 8 chain groups on their own host threads and streams); the roofline pass runs single-lane',
                       'newton_iters_mean': iters_total / chains_done, 'failed_chains': bad},
            'e2e': {'value': e2e_val, 'unit': UNIT, 'ms_per_step': ms_e2e / args.steps,
                    'h2d_bytes_per_step': int(B * n * N * 8 + B * (D + 1) * 8), 'd2h_bytes_per_step': int(B * (8 + 4 + 4))},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': roofline,
            'cpu_baseline': cpu,
            'cached_estimates_per_s': world * B * max(args.steps, 3) / (ms_cached * 1e-3),
            'apm_iters_per_s': {'value': world * B * apm_iters / t_apm, 'unit': 'chain-iterations/s',
                                'method': 'E-SS u + RD-SS theta, lock-step, %d chains/GPU, %d iterations, device RNG, asynchronous FULL rounds (worker thread + companion context for the CACHED rounds)' % (B, apm_iters),
                                'full_estimates_per_iter': float(apm_out['n_full'].mean() - 1) / apm_iters,
                                'cached_estimates_per_iter': float(apm_out['n_cached'].mean()) / apm_iters,
                                'failed_chains': int((apm_out['failed'] != 0).sum()), 'timing': 'host wall clock incl. the Python scheduler and the drain of the last iterations'},
            'diagnostics_gather': {'collective': 'nccl all_gather' if world > 1 else 'none (1 GPU)',
                                   'chains': int(all_logml.shape[0]), 'mean_logml': float(np.nanmean(all_logml))},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='apm_b200', choices=['apm_b200', 'reference'])
    ap.add_argument('--chains', type=int, default=0, help='chains per GPU (default 256)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--apm-iters', type=int, default=100, help='iterations of the batched ESS+RDSS sampler leg (chains drain at the end of a run: short runs understate the steady state)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == '__main__':
    sys.exit(main())
