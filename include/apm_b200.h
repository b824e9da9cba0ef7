/*
 * apm_b200.h -- C ABI of the B200-native pseudo-marginal likelihood engine (GP probit, Laplace
 * importance-sampling estimator).  This is the drop-in boundary for the hot path of
 * matt-graham/auxiliary-pm-mcmc: everything gpdemo.kernels / gpdemo.latent_posterior_approximations
 * / gpdemo.estimators compute, batched over independent chains, in fp64 on one GPU.
 *
 * Conventions
 *  - plain C, no torch / C++ types.  All functions return an apm_status (0 = ok) and never abort the
 *    process; apm_last_error() gives a message for the calling thread's last failure.
 *  - "host" pointers are ordinary CPU memory (pinned or pageable).  Bulk inputs that may already be
 *    resident in HBM (the auxiliary normals u) carry an explicit *_on_device flag; a device pointer is
 *    what torch.Tensor.data_ptr() returns for a CUDA tensor.
 *  - all matrices are row-major (numpy C order), fp64.  Targets y are +1/-1.
 *  - a "chain" is one independent Markov chain; B chains are processed in lock-step by one call.
 *  - a "slot" is a device-resident cache of (chol K, chol C, f_post, log-dets) for one theta, i.e. the
 *    reference's `cached_results` tuple (gpdemo/estimators.py:171-186).  Samplers keep <= 2 per chain.
 *  - per-chain status codes (chain_status[]) mirror the reference's exceptions:
 *      0 ok
 *      1 chol(K) failed            -> numpy.linalg.LinAlgError          (gpdemo/estimators.py:206)
 *      2 Newton > max_iters        -> MaximumIterationsExceededError    (lpa.py:100-102)
 *      3 chol(C) failed            -> InvalidCovarianceMatrixError      (gpdemo/estimators.py:208-215)
 *      4 non-finite input/result   -> ValueError (scipy check_finite)
 *      5 chol(B) failed            -> numpy.linalg.LinAlgError          (lpa.py:92)
 */
#ifndef APM_B200_H
#define APM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct apm_ctx apm_ctx;

typedef enum {
    APM_OK = 0,
    APM_ERR_INVALID = 1,   /* bad argument */
    APM_ERR_CUDA = 2,      /* CUDA runtime error (message in apm_last_error) */
    APM_ERR_NOMEM = 3,     /* device allocation failed */
    APM_ERR_NOGPU = 4      /* no CUDA device visible: there is NO CPU fallback */
} apm_status;

typedef enum {
    APM_KERNEL_ISO = 0,    /* gpdemo/kernels.pyx:12-49  theta = [log sigma, log tau]            */
    APM_KERNEL_ARD = 1     /* gpdemo/kernels.pyx:52-90  theta = [log sigma, log tau_1..tau_D]   */
} apm_kernel_kind;

enum { APM_CHAIN_OK = 0, APM_CHAIN_CHOL_K = 1, APM_CHAIN_NEWTON_MAXIT = 2, APM_CHAIN_CHOL_C = 3,
       APM_CHAIN_NONFINITE = 4, APM_CHAIN_CHOL_B = 5 };

/* library / build info: returns a static string "apm_b200 <version> sm_100a ..." */
const char* apm_version(void);
/* message for the last failing call on this thread */
const char* apm_last_error(void);

/*
 * Create an engine for one data set on one GPU.
 * Replaces the constructor state of gpdemo/estimators.py:112-144 (X, y, scratch K) for
 * `max_chains` chains at once.  X (n x D) and y (n) are HOST pointers, copied to the device.
 * n_slots cache slots are allocated (>= 2*max_chains for APM samplers); max_nimp is the largest
 * importance-sample count N that will be used.  device = CUDA ordinal.
 */
int apm_create(const double* X, const double* y, int n, int D, int kernel_kind, double epsilon,
               int max_chains, int n_slots, int max_nimp, int device, apm_ctx** out);
int apm_destroy(apm_ctx* ctx);

/* Companion context: streams, workspaces and pinned staging of its own on the parent's data set and device, but its
 * cache slots ARE the parent's.  apm_estimate_cached / apm_estimate_cached_weights on the companion may run while
 * another host thread is inside apm_estimate_full on the parent, provided the two calls touch different slots (the
 * reference's samplers hold a current and a proposed cache per chain: smp.py:401-417).  Only the cached estimates are
 * allowed on it; destroy it before the parent.  Used by apm_b200.batched's asynchronous scheduler. */
int apm_create_companion(apm_ctx* parent, int max_chains, int max_nimp, apm_ctx** out);

/* Run all work of this context on the given CUDA stream (cudaStream_t as an integer, e.g.
 * torch.cuda.current_stream().cuda_stream).  0 = the legacy default stream. */
int apm_set_stream(apm_ctx* ctx, uint64_t cuda_stream);
/* Block the host until everything queued by this context has finished. */
int apm_synchronize(apm_ctx* ctx);

/* FULL estimates overlap work that is off the critical path (chol(K) on an auxiliary stream, H2D of u on a copy
 * stream).  enable = 0 serialises everything on the context's stream (used for unperturbed per-kernel timing). */
int apm_set_overlap(apm_ctx* ctx, int enable);

/* Laplace/Newton controls, defaults as lpa.py:22-23: diff_f_tol = 1e-4, max_iters = 1000. */
int apm_set_newton(apm_ctx* ctx, double diff_f_tol, int max_iters);

/*
 * Posterior approximation used by apm_estimate_full: kind 0 = Laplace (the reference's
 * gpdemo.latent_posterior_approximations.laplace_approximation, the default), kind 1 = expectation propagation.
 * EP is an EXTENSION: the project brief names it, the reference does not contain it, so it
 * plugs in where the reference would take it -- the post_approx_func(K, y) -> (f_post, C, cubic_ops) argument of
 * LogMarginalLikelihoodApproxPosteriorISEstimator (gpdemo/estimators.py:126-139).  Algorithm: GPML Alg. 3.5 with
 * all sites updated per sweep ("parallel EP"), stopped when the largest change of a site parameter is < ep_tol;
 * ep_damping in (0, 1]; the CPU restatement it is checked against is oracle/apm_oracle.py:ep_approximation.
 * The ep_* arguments are ignored for kind 0.
 */
int apm_set_approximation(apm_ctx* ctx, int kind, double ep_tol, int ep_max_iters, double ep_damping);

/* Problem geometry: n, D, padded n, number of theta components, slots, max chains. */
int apm_get_info(apm_ctx* ctx, int* n, int* D, int* n_pad, int* n_theta, int* n_slots,
                 int* max_chains, int* max_nimp);

/*
 * K(theta) builders -- replace gpdemo.kernels.{isotropic,diagonal}_squared_exponential_kernel
 * (gpdemo/kernels.pyx:12-49, 52-90).  theta: HOST [B][n_theta].  K_out: [B][n][n] dense symmetric,
 * written to host memory (K_on_device = 0) or device memory (1).  kernel_kind / epsilon override the
 * context's (pass kernel_kind < 0 / epsilon < 0 to keep them).
 */
int apm_kernel_build(apm_ctx* ctx, const double* theta, int B, int kernel_kind, double epsilon,
                     double* K_out, int K_on_device);

/*
 * Gradients of the covariance with respect to the (log-)hyper-parameters, dK/dtheta_p for p = 0..n_theta-1.
 * EXTENSION: named by the project brief next to the K build; the reference has no gradient code (no
 * gradient-based update exists in auxpm), so there is no reference interface this replaces.
 *   ARD: dK_ij/dtheta_0 = K_ij - epsilon [i==j];  dK_ij/dtheta_{k+1} = K_ij ((x_ik - x_jk) / exp(theta_{k+1}))^2
 *   ISO: dK_ij/dtheta_0 as above;                  dK_ij/dtheta_1 = K_ij |x_i - x_j|^2 / exp(theta_1)^2
 * theta: HOST [B][n_theta].  dK_out: [B][n_theta][n][n] dense, host (dK_on_device = 0) or device (1).
 */
int apm_kernel_grad(apm_ctx* ctx, const double* theta, int B, int kernel_kind, double* dK_out, int dK_on_device);

/*
 * Laplace approximation for caller-supplied covariance matrices -- replaces
 * gpdemo.latent_posterior_approximations.laplace_approximation (lpa.py:22-124).
 * K: [B][n][n] host or device.  Outputs (any may be NULL): f_out HOST [B][n] posterior mode;
 * C_out [B][n][n] posterior covariance (host or device, only if calc_cov); lml_out HOST [B]
 * (only if calc_lml); cubic_ops_out HOST [B] = iterations (+1 if calc_cov), lpa.py:113-124;
 * chain_status HOST [B].
 */
int apm_laplace(apm_ctx* ctx, const double* K, int K_on_device, int B, int calc_cov, int calc_lml,
                double* f_out, double* C_out, int C_on_device, double* lml_out, int* cubic_ops_out,
                int* chain_status);

/*
 * EP posterior approximation for given covariance matrices (extension, see apm_set_approximation; uses the
 * context's ep_tol / ep_max_iters / ep_damping).  K: [B][n][n]; f_out HOST [B][n] posterior mean; C_out [B][n][n]
 * posterior covariance (if calc_cov); nu_out / tau_out HOST [B][n] site parameters (may be NULL);
 * cubic_ops_out = EP iterations (+1 with calc_cov); chain_status as apm_laplace (2 = iteration limit).
 */
int apm_ep(apm_ctx* ctx, const double* K, int K_on_device, int B, int calc_cov, double* f_out, double* C_out,
           int C_on_device, double* nu_out, double* tau_out, int* cubic_ops_out, int* chain_status);

/*
 * FULL importance-sampling estimate -- replaces
 * LogMarginalLikelihoodApproxPosteriorISEstimator.__call__ with cached_results=None
 * (gpdemo/estimators.py:203-241): K(theta) -> chol K -> Laplace (Newton + covariance) -> chol C ->
 * f = mu + L_C u -> probit log-lik, log p(f), log q(f) -> logsumexp - log N.
 *   theta   HOST [B][n_theta]
 *   u       [B][n][N] (reference layout, element (i,s) at i*N+s), host or device
 *   slots   HOST [B]: cache slot written for chain b (the returned cached_results)
 *   logml_out HOST [B]; cubic_ops_out HOST [B] = newton iters + 1 + 2 (est.py:217), may be NULL;
 *   chain_status HOST [B].
 * Chains with a non-zero status get logml = NaN and leave their slot invalid.
 * Execution: one host thread; the Newton rounds are queued under device-side masks (one host round trip per call on
 * typical data), chol(K) and the upload of a host u overlap the mode search on auxiliary streams.  A chain's result
 * does not depend on the other chains of the batch.
 */
int apm_estimate_full(apm_ctx* ctx, const double* theta, const double* u, int u_on_device, int N,
                      int B, const int* slots, double* logml_out, int* cubic_ops_out,
                      int* chain_status);

/*
 * CACHED estimate (u changed only) -- gpdemo/estimators.py:218-241 with cached_results given.
 * O(n^2 N) per chain: reads the slot's chol K / chol C / f_post.
 */
int apm_estimate_cached(apm_ctx* ctx, const int* slots, const double* u, int u_on_device, int N,
                        int B, double* logml_out, int* chain_status);

/*
 * As the two calls above but also returns the per-importance-sample log weights
 * (est.py:238-239, before the logsumexp) into logw_out HOST [B][N]; used by tests.
 */
int apm_estimate_cached_weights(apm_ctx* ctx, const int* slots, const double* u, int u_on_device,
                                int N, int B, double* logw_out);

/*
 * Deterministic Laplace log-marginal-likelihood -- replaces
 * LogMarginalLikelihoodLaplaceEstimator.__call__ (gpdemo/estimators.py:65-82):
 * K(theta) -> Newton (calc_cov=False, calc_lml=True).  cubic_ops_out = newton iterations.
 */
int apm_laplace_lml(apm_ctx* ctx, const double* theta, int B, double* lml_out, int* cubic_ops_out,
                    int* chain_status);

/*
 * Prior Monte-Carlo estimate -- replaces LogMarginalLikelihoodPriorMCEstimator.__call__
 * (gpdemo/estimators.py:297-325): f = L_K u, logsumexp_s sum_i log Phi(y_i f_is) - log N.
 * If theta != NULL the slot's chol K is (re)built first (one cubic op), else the slot is reused.
 */
int apm_estimate_prior_mc(apm_ctx* ctx, const double* theta, const int* slots, const double* u,
                          int u_on_device, int N, int B, double* logml_out, int* chain_status);

/*
 * Export one slot's cache to HOST memory in the reference's format (est.py:240-241):
 * K_chol, C_chol [n][n] lower triangular with zeroed upper triangle, f_post [n].  NULL = skip.
 */
int apm_slot_export(apm_ctx* ctx, int slot, double* K_chol, double* C_chol, double* f_post,
                    double* logdets2);
/* Import a cache from host arrays (a cached_results tuple produced elsewhere) into a slot. */
int apm_slot_import(apm_ctx* ctx, int slot, const double* K_chol, const double* C_chol,
                    const double* f_post);
/* Build a slot from caller-supplied dense matrices: chol(K) and chol(C) are computed on the device and
 * stored with f_post -- the cache a foreign `post_approx_func(K, y) -> (f_post, C, ops)` plug-in
 * (gpdemo/estimators.py:126-139, 206-209) needs.  K, C: [n][n] host or device; f_post HOST [n].
 * chain_status HOST [1] receives 0 / 1 (chol K failed) / 3 (chol C failed). */
int apm_slot_factor(apm_ctx* ctx, int slot, const double* K, const double* C, int on_device,
                    const double* f_post, int* chain_status);
/* Copy slot src -> dst on the device (accepting a proposal: cached_res_curr = cached_res_prop,
 * auxpm/samplers.py:413-417), for B (src,dst) pairs given as HOST arrays. */
int apm_slot_copy(apm_ctx* ctx, const int* src, const int* dst, int B);

/* Per-kernel device timing: when enabled every launch on the context's stream is bracketed by CUDA
 * events; apm_profile_read returns, per kernel family, the accumulated milliseconds and launch counts
 * (names: max_entries x 32 chars).  Returns the number of entries, < 0 on error.  bench.py uses this for the
 * live roofline numbers. */
int apm_profile(apm_ctx* ctx, int enable);
int apm_profile_read(apm_ctx* ctx, int max_entries, char* names, double* ms, int64_t* counts, int reset);

/* Number of kernels this context has launched since creation / the last reset (bench.py reports
 * it as gpu_launches). */
int64_t apm_launch_count(apm_ctx* ctx, int reset);

/* Work the DMMA kernel families actually executed since creation / the last reset, in units of n^3/3 flops per chain:
 * out[0] = chain-Choleskys factored by k_chol_flow (masked-out chains and the factorisations the hybrid Newton
 * iteration skips are not counted), out[1] = M' = I + L_K^T W L_K builds accumulated inside k_chol_flow<true>.  bench.py's per-kernel roofline
 * uses these instead of the reference's nominal operation count (lpa.py:92, 111-112; est.py:206, 209). */
int apm_work_count(apm_ctx* ctx, int64_t* out, int reset);

/* Tuning aid: average milliseconds of one batched Cholesky of the context's current K matrices (B chains,
 * after apm_kernel_build) into slots 0..B-1.  mode 0 = default path, 1 = per-step launches. */
int apm_dev_chol_bench(apm_ctx* ctx, int B, int reps, int mode, double* ms_out);

/*
 * Native batched sampler -- the lock-step drivers of SURVEY.md section 8 f-1: B independent chains of one of the
 * reference's composite samplers (auxpm/samplers.py: APM MI+MH :346-418, APM ESS+MH :515-587, APM MI+RDSS :658-730,
 * APM ESS+RDSS :800-841, PM-MH :223-262; updates of auxpm/mcmc_updates.py :117-160, :284-303, :373-400, :481-519) advanced
 * together on the engine `ctx`.  The auxiliary normals are generated on the device (Philox4x32-10 keyed by the chain's
 * seed) directly in the layout the estimator reads and never leave it; proposals, the ellipse u cos(phi) + v sin(phi)
 * and the accepted state are device buffers; the per-chain scalars (theta, brackets, accept / reject) are host state.
 * The log prior added to every estimate is the notebooks' log-Gamma prior (gpdemo/utils.py:39-59, nb cell 12):
 * prior_ab HOST [n_theta][2] = shape a and rate b of every theta component.
 *   method       0 MI+MH, 1 ESS+MH, 2 MI+RDSS, 3 ESS+RDSS, 4 PM-MH
 *   seeds        HOST [n_chains]: a chain's trace depends on its seed only (not on the batch, the schedule or the GPU count)
 *   prop_scales  HOST [n_theta] random-walk scales of the MH theta-update (methods 0, 1, 4), else NULL
 *   slice_width  w of the random-direction slice update (methods 2, 3; max_steps_out = 0 as in the notebooks)
 * ctx must have max_chains >= n_chains, n_slots >= 2 n_chains (current / proposed cache per chain) and max_nimp >= n_imp;
 * it must not be used by other calls while apm_sampler_run is running.
 */
typedef struct apm_sampler apm_sampler;
int apm_sampler_create(apm_ctx* ctx, int method, int n_chains, int n_imp, const uint64_t* seeds, const double* prior_ab,
                       const double* prop_scales, double slice_width, int max_slice_iters, apm_sampler** out);
/* n_sample states per chain (the first is theta_init, as get_samples of the reference returns them).
 * theta_init HOST [n_chains][n_theta]; thetas_out HOST [n_chains][n_sample][n_theta] (NaN after a chain failed);
 * counts_out HOST [n_chains][6]: rejected u-updates, rejected theta-updates, n_cubic_ops, FULL estimates, CACHED
 * estimates, failure status (0 = ok, else the chain_status code that stopped the chain). */
int apm_sampler_run(apm_sampler* s, const double* theta_init, int n_sample, double* thetas_out, int64_t* counts_out);
/* Scheduling diagnostics of the last run: out8 = FULL calls, chains in them, CACHED calls, chains in them, summed
 * durations of the FULL calls [s], seconds in total, scheduler rounds, seconds with at least one FULL call in flight.
 * Environment (read by apm_sampler_create): APM_SAMPLER_BATCH_FRAC (share of the waiting chains that must ask for a FULL
 * estimate before a call is started, 0.5), APM_SAMPLER_JOBS=2 (a second FULL call may start beside one in flight),
 * APM_SAMPLER_MIN_SECOND (smallest such second call). */
int apm_sampler_stats(apm_sampler* s, double* out8);
int apm_sampler_destroy(apm_sampler* s);

/* Micro-benchmarks used by bench.py to measure the fp64 roofline denominators on the box:
 * kind 0: DMMA m8n8k4 issue peak, 1: DFMA peak.  Returns TFLOP/s in *tflops. */
int apm_measure_fp64_peak(int device, int kind, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* APM_B200_H */
