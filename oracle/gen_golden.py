"""Generate tests/golden/*.npz by running the UNMODIFIED reference (via oracle/ref_loader.py).

Run in the build container only (needs /root/reference):
    OPENBLAS_NUM_THREADS=1 python oracle/gen_golden.py

The reference has no tests or fixtures of its own, so these vectors -- outputs of the
reference itself on seeded synthetic inputs -- are what pins both the oracle restatement
(oracle/apm_oracle.py) and the CUDA path.  Inputs are stored in the fixture (X, y, theta) or regenerated
from `numpy.random.RandomState(seed)` legacy streams, which are frozen across numpy versions.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
warnings.simplefilter('ignore', SyntaxWarning)
import ref_loader  # noqa: E402
from apm_b200 import synth  # noqa: E402  (data generation only)

OUT = os.path.join(ROOT, 'tests', 'golden')
EPS = 1e-8


def kernels_of(ref, kind):
    if kind == 'iso':
        return lambda K, X, th: ref.kernels.isotropic_squared_exponential_kernel(K, X, th, EPS)
    return lambda K, X, th: ref.kernels.diagonal_squared_exponential_kernel(K, X, th, EPS)


def gen_kernels(ref):
    rs = np.random.RandomState(101)
    out = {}
    for tag, (n, D) in {'a': (13, 3), 'b': (70, 9)}.items():
        X = rs.normal(size=(n, D))
        th_iso = rs.normal(size=(3, 2)) * 0.7
        th_ard = rs.normal(size=(3, D + 1)) * 0.7
        K_iso = np.empty((3, n, n))
        K_ard = np.empty((3, n, n))
        for t in range(3):
            ref.kernels.isotropic_squared_exponential_kernel(K_iso[t], X, th_iso[t], EPS)
            ref.kernels.diagonal_squared_exponential_kernel(K_ard[t], X, th_ard[t], 1e-6)
        out.update({'X_' + tag: X, 'th_iso_' + tag: th_iso, 'th_ard_' + tag: th_ard,
                    'K_iso_' + tag: K_iso, 'K_ard_' + tag: K_ard})
    out['eps_iso'] = EPS
    out['eps_ard'] = 1e-6
    np.savez_compressed(os.path.join(OUT, 'kernels.npz'), **out)


def gen_laplace(ref):
    out = {}
    for tag, (n, D, seed) in {'a': (40, 3, 5), 'b': (150, 6, 6)}.items():
        X, y, th = synth.make_dataset(n, D, seed)
        K = np.empty((n, n))
        ref.kernels.diagonal_squared_exponential_kernel(K, X, th, EPS)
        f, C, lml, ops = ref.lpa.laplace_approximation(K, y, calc_cov=True, calc_lml=True)
        f2, lml2, ops2 = ref.lpa.laplace_approximation(K, y, calc_cov=False, calc_lml=True)
        assert np.array_equal(f, f2)
        out.update({'K_' + tag: K, 'y_' + tag: y, 'f_' + tag: f, 'C_' + tag: C, 'lml_' + tag: lml,
                    'ops_cov_' + tag: ops, 'ops_nocov_' + tag: ops2})
    np.savez_compressed(os.path.join(OUT, 'laplace.npz'), **out)


def gen_estimator(ref, name, n, D, kind, Ns, n_theta, seed, keep_mats):
    X, y, th_true = synth.make_dataset(n, D, seed)
    P = D + 1 if kind == 'ard' else 2
    rs = np.random.RandomState(seed + 1000)
    thetas = th_true[:P][None] + 0.3 * rs.normal(size=(n_theta, P))
    kf = kernels_of(ref, kind)
    out = dict(X=X, y=y, thetas=thetas, Ns=np.array(Ns), kind=kind, eps=EPS)
    for t in range(n_theta):
        for N in Ns:
            est = ref.est.LogMarginalLikelihoodApproxPosteriorISEstimator(X, y, kf, ref.lpa.laplace_approximation)
            u1 = np.random.RandomState(7000 + 10 * t + N).normal(size=(n, N))
            u2 = np.random.RandomState(8000 + 10 * t + N).normal(size=(n, N))
            full, cache = est(u1, thetas[t])
            cached, _ = est(u2, None, cache)
            key = 't%d_N%d_' % (t, N)
            out[key + 'full'] = full
            # noise floor of this case: how far the REFERENCE's own answer moves when every K entry is
            # perturbed by <= 1 ulp (device exp vs libm exp differ by that much).  ~cond(K) * 1e-16.
            def kf_ulp(K_out, X_, th_):
                kf(K_out, X_, th_)
                E = np.random.RandomState(1).uniform(-1, 1, size=K_out.shape) * 1.1e-16
                K_out *= 1 + (E + E.T) / 2
            est_p = ref.est.LogMarginalLikelihoodApproxPosteriorISEstimator(X, y, kf_ulp, ref.lpa.laplace_approximation)
            out[key + 'ulp_sens'] = abs(est_p(u1, thetas[t])[0] - full)
            out[key + 'cached'] = cached
            out[key + 'cubic_ops'] = est.n_cubic_ops
        # per-theta cache summaries (from the last N; caches do not depend on u)
        K_tmp = np.empty((n, n))
        kf(K_tmp, X, thetas[t])
        out['t%d_condK' % t] = np.linalg.cond(K_tmp)
        out['t%d_f_post' % t] = cache[2]
        out['t%d_diagK' % t] = cache[0].diagonal().copy()
        out['t%d_diagC' % t] = cache[1].diagonal().copy()
        if keep_mats:
            out['t%d_K_chol' % t] = cache[0]
            out['t%d_C_chol' % t] = cache[1]
        lap = ref.est.LogMarginalLikelihoodLaplaceEstimator(X, y, kf)
        out['t%d_laplace_lml' % t] = lap(thetas[t])
        out['t%d_laplace_ops' % t] = lap.n_cubic_ops
        pm = ref.est.LogMarginalLikelihoodPriorMCEstimator(X, y, kf)
        u3 = np.random.RandomState(9000 + t).normal(size=(n, Ns[-1]))
        out['t%d_prior_mc' % t], _ = pm(u3, thetas[t])
    np.savez_compressed(os.path.join(OUT, 'estimator_%s.npz' % name), **out)


def build_sampler(ref, method, X, y, N, prng, est_holder):
    """The wiring of the reference notebooks (cells 8-12) on a given data set."""
    D = X.shape[1]
    prior = synth.prior_params(D)
    lg = ref.utils.log_gamma_log_pdf
    kf = kernels_of(ref, 'iso')
    ml = ref.est.LogMarginalLikelihoodApproxPosteriorISEstimator(X, y, kf, ref.lpa.laplace_approximation)
    est_holder.append(ml)

    def log_prior(theta):
        return lg(theta[0], prior['a_sigma'], prior['b_sigma']) + lg(theta[1], prior['a_tau'], prior['b_tau'])

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
        return ref.smp.APMMetIndPlusMHSampler(log_f_estimator, log_prop_density, prop_sampler, scales, u_sampler, prng)
    if method == 'ess+mh':
        return ref.smp.APMEllSSPlusMHSampler(log_f_estimator, log_prop_density, prop_sampler, scales, u_sampler, prng)
    if method == 'mi+rdss':
        return ref.smp.APMMetIndPlusRandDirSliceSampler(log_f_estimator, u_sampler, prng, dir_and_w, 0)
    if method == 'ess+rdss':
        return ref.smp.APMEllSSPlusRandDirSliceSampler(log_f_estimator, u_sampler, prng, dir_and_w, 0)
    if method == 'pmmh':
        main = lambda th: ml(prng.normal(size=(y.shape[0], N)), th)[0] + log_prior(th)  # noqa: E731
        return ref.smp.PMMHSampler(main, log_prop_density, prop_sampler, scales, prng)
    raise ValueError(method)


def gen_samplers(ref):
    n, D, n_iter = 60, 3, 1000
    X, y, _ = synth.make_dataset(n, D, seed=21)
    out = dict(X=X, y=y, n_iter=n_iter)
    for method in ['mi+mh', 'ess+mh', 'mi+rdss', 'ess+rdss', 'pmmh']:
        for N in (1, 4):
            prng = np.random.RandomState()
            holder = []
            smp = build_sampler(ref, method, X, y, N, prng, holder)
            prng.seed(1000 + N)
            theta_init = synth.draw_theta_prior(prng, D, ard=False)
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                res = smp.get_samples(theta_init, n_iter)
            thetas = res[0] if isinstance(res, tuple) else res
            n_rej = np.atleast_1d(res[1]) if isinstance(res, tuple) else np.zeros(0)
            key = '%s_N%d_' % (method, N)
            out[key + 'thetas'] = thetas
            out[key + 'n_reject'] = np.asarray(n_rej, dtype=np.int64)
            out[key + 'cubic_ops'] = holder[0].n_cubic_ops
            print(method, N, 'final theta', thetas[-1], 'rej', n_rej, 'ops', holder[0].n_cubic_ops)
    # adaptive MH run (smp.py:68-156 + utils.py:62-83) for the MI+MH sampler
    prng = np.random.RandomState()
    holder = []
    smp = build_sampler(ref, 'mi+mh', X, y, 1, prng, holder)
    prng.seed(4242)
    theta_init = synth.draw_theta_prior(prng, D, ard=False)
    smp.prop_scales = np.array([0.5, 0.5])
    th, sc, acc = smp.adaptive_run(theta_init, 25, 8, 0.15, 0.30, ref.utils.adapt_factor_func)
    out['adapt_thetas'] = th
    out['adapt_scales'] = sc
    out['adapt_accept'] = acc
    np.savez_compressed(os.path.join(OUT, 'samplers.npz'), **out)


def gen_samplers_fullsize(ref):
    """Chains of the unmodified reference at the headline shapes (BASELINE configs 1-2: pima n=768 D=8 and breast n=682
    D=9, N_imp = 64, isotropic kernel as in the notebooks): accept/reject parity where the benchmark is quoted."""
    out = {}
    for tag, n, D, seed, cases in (('pima', 768, 8, 0, (('ess+rdss', 50), ('mi+mh', 50), ('pmmh', 40))),
                                   ('breast', 682, 9, 1, (('ess+rdss', 50),))):
        X, y, _ = synth.make_dataset(n, D, seed=seed)
        out[tag + '_X'], out[tag + '_y'] = X, y
        for method, n_iter in cases:
            N = 64
            prng = np.random.RandomState()
            holder = []
            smp = build_sampler(ref, method, X, y, N, prng, holder)
            prng.seed(1000 + N)
            theta_init = synth.draw_theta_prior(prng, D, ard=False)
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                res = smp.get_samples(theta_init, n_iter)
            thetas = res[0] if isinstance(res, tuple) else res
            n_rej = np.atleast_1d(res[1]) if isinstance(res, tuple) else np.zeros(0)
            key = '%s_%s_N%d_' % (tag, method, N)
            out[key + 'thetas'] = thetas
            out[key + 'n_reject'] = np.asarray(n_rej, dtype=np.int64)
            out[key + 'cubic_ops'] = holder[0].n_cubic_ops
            print(tag, method, N, 'final theta', thetas[-1], 'rej', n_rej, 'ops', holder[0].n_cubic_ops, flush=True)
    np.savez_compressed(os.path.join(OUT, 'samplers_fullsize.npz'), **out)


def gen_samplers_long(ref, which=None):
    """1000-iteration chains of the unmodified reference at the headline shape (pima n=768, D=8, N_imp=64, isotropic
    kernel as in the notebooks): the north-star's accept/reject bar ("first 1000 iterations") where the benchmark is quoted.
    One file per sampler (tests/golden/samplers_long_<method>.npz) so the two can be generated in parallel
    (`--only samplers_long ess+rdss`); ~10-20 min of CPU each."""
    n, D, N, n_iter = 768, 8, 64, 1000
    X, y, _ = synth.make_dataset(n, D, seed=0)
    for method in ([which] if which else ['ess+rdss', 'mi+mh']):
        prng = np.random.RandomState()
        holder = []
        smp = build_sampler(ref, method, X, y, N, prng, holder)
        prng.seed(1000 + N)
        theta_init = synth.draw_theta_prior(prng, D, ard=False)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            res = smp.get_samples(theta_init, n_iter)
        thetas = res[0] if isinstance(res, tuple) else res
        n_rej = np.atleast_1d(res[1]) if isinstance(res, tuple) else np.zeros(0)
        out = dict(n=n, D=D, N=N, n_iter=n_iter, data_seed=0, chain_seed=1000 + N, thetas=thetas,
                   n_reject=np.asarray(n_rej, dtype=np.int64), cubic_ops=holder[0].n_cubic_ops)
        print('long', method, 'final theta', thetas[-1], 'rej', n_rej, 'ops', holder[0].n_cubic_ops, flush=True)
        np.savez_compressed(os.path.join(OUT, 'samplers_long_%s.npz' % method.replace('+', '_')), **out)


def gen_estimator_iters(ref):
    """Headline-shape (pima n=768, D=8, ARD, N_imp=64) estimator goldens chosen so that the Newton loop takes
    I = 3, 4, 5, 6 iterations (cubic_ops 6..9): the hybrid-Newton prediction is exercised both ways against the
    REFERENCE (not against the repo's own APM_NO_HYBRID_NEWTON build).  Candidates are scanned with the cheap
    Laplace-only estimator; one theta per iteration count is kept."""
    n, D, N = 768, 8, 64
    X, y, th_true = synth.make_dataset(n, D, 0)
    kf = kernels_of(ref, 'ard')
    rs = np.random.RandomState(4711)
    found = {}
    cands = []
    for ls in (-2.5, -1.5, -0.5, 0.5, 1.5, 2.5, 3.5, 4.5):
        for lt in (0.0, 0.7, 1.4):
            cands.append(np.r_[ls, lt + 0.2 * rs.normal(size=D)])
    for th in cands:
        lap = ref.est.LogMarginalLikelihoodLaplaceEstimator(X, y, kf)
        try:
            lap(th)
        except Exception as e:   # chol failure at extreme theta: not a candidate
            print('skip', th[:2], type(e).__name__, flush=True)
            continue
        it = lap.n_cubic_ops
        print('theta0 %.1f tau~%.1f -> I = %d' % (th[0], th[1], it), flush=True)
        found.setdefault(it, []).append(th)
        if all(len(found.get(k, [])) >= 1 for k in (3, 4, 5, 6)) and len(found.get(5, [])) >= 2:
            break
    thetas, iters = [], []
    for it in sorted(found):
        for th in found[it][:2 if it in (3, 5, 6) else 1]:
            thetas.append(th)
            iters.append(it)
    thetas = np.array(thetas)
    out = dict(X=X, y=y, thetas=thetas, newton_iters=np.array(iters), N=N, kind='ard', eps=EPS)
    for t, th in enumerate(thetas):
        est = ref.est.LogMarginalLikelihoodApproxPosteriorISEstimator(X, y, kf, ref.lpa.laplace_approximation)
        u1 = np.random.RandomState(7100 + t).normal(size=(n, N))
        u2 = np.random.RandomState(8100 + t).normal(size=(n, N))
        full, cache = est(u1, th)
        cached, _ = est(u2, None, cache)

        def kf_ulp(K_out, X_, th_):
            kf(K_out, X_, th_)
            E = np.random.RandomState(1).uniform(-1, 1, size=K_out.shape) * 1.1e-16
            K_out *= 1 + (E + E.T) / 2
        est_p = ref.est.LogMarginalLikelihoodApproxPosteriorISEstimator(X, y, kf_ulp, ref.lpa.laplace_approximation)
        key = 't%d_' % t
        out[key + 'full'] = full
        out[key + 'cached'] = cached
        out[key + 'cubic_ops'] = est.n_cubic_ops
        out[key + 'ulp_sens'] = abs(est_p(u1, th)[0] - full)
        out[key + 'f_post'] = cache[2]
        K_tmp = np.empty((n, n))
        kf(K_tmp, X, th)
        out[key + 'condK'] = np.linalg.cond(K_tmp)
        print('iters golden', t, 'I', iters[t], 'ops', est.n_cubic_ops, 'full', full, 'sens', out[key + 'ulp_sens'],
              'condK %.2e' % out[key + 'condK'], flush=True)
    np.savez_compressed(os.path.join(OUT, 'estimator_pima_iters.npz'), **out)


def gen_utils(ref):
    xs = np.linspace(-3, 3, 13)
    np.savez_compressed(
        os.path.join(OUT, 'utils.npz'), xs=xs,
        lg_11_01=ref.utils.log_gamma_log_pdf(xs, 1.1, 0.1), lg_1_03=ref.utils.log_gamma_log_pdf(xs, 1., 0.3),
        adapt=np.array([ref.utils.adapt_factor_func(b, 20) for b in range(20)]))


def gen_run_artefacts(ref):
    """Files written by the reference's own save_run / save_adaptive_run (gpdemo/utils.py:108-208) for fixed arguments; the
    time stamp prefix of the file names is stripped.  tests/test_samplers_host.py compares the product's files with them."""
    import glob
    import shutil
    import tempfile
    out = os.path.join(OUT, 'run_artefacts')
    os.makedirs(out, exist_ok=True)
    thetas = np.arange(12.).reshape(6, 2)
    with tempfile.TemporaryDirectory() as tmp:
        ref.utils.save_run(tmp, 'apm_test', thetas, (3, 4), 77, 1.5, {'n_imp': 4, 'a': [1, 2], 'tag': 'x'})
        ref.utils.save_run(tmp, 'pmmh_test', thetas, 5, 9, 2.5, {'seed': 1})
        ref.utils.save_adaptive_run(tmp, 'ad_test', thetas, thetas[:3], np.ones(3) * 0.25, thetas + 1, (1, 2), 9, 2.5,
                                    {'n_batch': 3, 'batch_size': 2})
        for f in sorted(glob.glob(os.path.join(tmp, '*'))):
            name = os.path.basename(f)[len('YYYY_mm_dd_HH_MM_SS_'):]
            shutil.copy(f, os.path.join(out, 'ref_' + name))
    print('run artefacts:', sorted(os.listdir(out)))


if __name__ == '__main__':
    if os.environ.get('OPENBLAS_NUM_THREADS') != '1':
        print('warning: run with OPENBLAS_NUM_THREADS=1 for a deterministic oracle', file=sys.stderr)
    ref = ref_loader.load_reference()
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 2 and sys.argv[1] == '--only':          # regenerate one fixture file, leave the others alone
        globals()['gen_' + sys.argv[2]](ref, *sys.argv[3:])
        sys.exit(0)
    gen_utils(ref)
    gen_run_artefacts(ref)
    gen_kernels(ref)
    gen_laplace(ref)
    gen_estimator(ref, 'small_ard', 100, 4, 'ard', (1, 8), 3, 31, True)
    gen_estimator(ref, 'small_iso', 90, 3, 'iso', (1, 70), 3, 32, True)
    gen_estimator(ref, 'pima_ard', 768, 8, 'ard', (1, 64), 3, 0, False)
    gen_estimator(ref, 'pima_iso', 768, 8, 'iso', (1, 64), 2, 0, False)
    gen_estimator(ref, 'breast_ard', 682, 9, 'ard', (64,), 2, 1, False)
    gen_samplers(ref)
    gen_samplers_fullsize(ref)
    gen_estimator_iters(ref)
    gen_samplers_long(ref)
    print('golden vectors written to', OUT)
