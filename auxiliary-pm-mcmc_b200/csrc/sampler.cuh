// sampler.cuh -- native batched auxiliary pseudo-marginal sampler; included at the end of apm_capi.cu.
//
// What it replaces: the per-chain Python control flow of auxpm/samplers.py (APM MI+MH :346-418, ESS+MH :515-587,
// MI+RDSS :658-730, ESS+RDSS :800-841, PM-MH :223-262) and auxpm/mcmc_updates.py (MH :117-160, MI u-update :284-303,
// elliptical slice :373-400, random-direction slice :481-519) for B chains at once.  The chain logic (order of random
// draws, brackets, cache hand-over) is the one of apm_b200.batched's generators, which are pinned to the reference's
// chains; here it is a C++ state machine per chain, the auxiliary normals never leave the GPU and never exist in the
// reference layout:
//   * u, v and the proposals live as U^T [chain][Npad][np] -- the layout the importance-sampling tail reads -- so a
//     CACHED estimate needs no transpose and no copy of u: the Philox normals (k_sampler_normals) or the ellipse point
//     u cos(phi) + v sin(phi) (k_sampler_ellipse, mu.py:382) are written straight into the engine's U^T workspace;
//   * random numbers are counter-based (Philox4x32-10, keyed by the chain's seed; apm_b200/philox.py is the numpy mirror),
//     so a chain's trace depends on its seed only -- not on the batch it runs in, the scheduling order or the GPU count;
//   * scheduling is the asynchronous scheme of apm_b200.batched: FULL estimates (theta changed; O(n^3), ~10 ms per call)
//     run on a worker thread and their own stream while this thread keeps serving the CACHED estimates (u changed;
//     O(n^2 N)) of the chains that are in their u-update through a companion context on the same cache slots.
// Accept / reject needs one scalar per chain and call (the estimate the C ABI returns anyway): it is decided on the host
// from the chain's scalar Philox stream; everything that is O(nN) stays on the device.
#pragma once
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <mutex>
#include <thread>

namespace apm {

// ---- Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) ---------------
struct Philox4 { uint32_t v[4]; };
__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox4 o; o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}
__host__ __device__ inline double philox_u01(uint32_t lo, uint32_t hi) {
    const uint64_t x = (((uint64_t)hi << 32) | lo) >> 11;
    return ((double)x + 0.5) * 1.1102230246251565e-16;   // 2^-53: uniform in (0, 1)
}
enum { PHILOX_SCALAR = 1, PHILOX_BULK = 2 };

// ---- device kernels -----------------------------------------------------------------------------------------------
// Standard normals of bulk draw draw_of[j] of chain chain_of[j] into the U^T block dst + dst_idx[j] * dst_bs
// ([Npad][np], zero padded): element (i, s) of the reference's (n, N) array u (nb cell 12: prng.normal(size=(n, N)))
// lands at [s][i].  grid (np / 128, Npad / 2, jobs): thread = data point i (coalesced rows), block row = sample pair.
__global__ void __launch_bounds__(128) k_sampler_normals(double* __restrict__ dst, long long dst_bs, const int* __restrict__ dst_idx,
                                                         const unsigned long long* __restrict__ seeds, const int* __restrict__ chain_of,
                                                         const unsigned* __restrict__ draw_of, int n, int N, int np) {
    const int j = blockIdx.z, q = blockIdx.y;
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= np) return;
    double z0 = 0.0, z1 = 0.0;
    if (i < n && 2 * q < N) {
        const unsigned long long seed = seeds[chain_of[j]];
        const unsigned long long idx = (unsigned long long)i * (unsigned)((N + 1) / 2) + (unsigned)q;
        const Philox4 r = philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), draw_of[j], PHILOX_BULK, (uint32_t)seed, (uint32_t)(seed >> 32));
        const double u1 = philox_u01(r.v[0], r.v[1]), u2 = philox_u01(r.v[2], r.v[3]);
        const double rad = sqrt(-2.0 * log(u1));
        double sn, cs;
        sincospi(2.0 * u2, &sn, &cs);
        z0 = rad * cs;
        if (2 * q + 1 < N) z1 = rad * sn;
    }
    double* d = dst + (long long)dst_idx[j] * dst_bs + (size_t)(2 * q) * np + i;
    d[0] = z0;
    d[np] = z1;
}

// dst[dst_idx[j]] = cs[j] * U[chain_of[j]] + sn[j] * V[chain_of[j]]   (mu.py:382), count doubles per block (even)
__global__ void __launch_bounds__(256) k_sampler_ellipse(double* __restrict__ dst, long long dst_bs, const int* __restrict__ dst_idx,
                                                         const double* __restrict__ U, const double* __restrict__ V, long long u_bs,
                                                         const int* __restrict__ chain_of, const double* __restrict__ cs,
                                                         const double* __restrict__ sn, long long count) {
    const int j = blockIdx.y;
    const long long e = 2 * ((long long)blockIdx.x * 256 + threadIdx.x);
    if (e >= count) return;
    const long long c = chain_of[j];
    const double2 u = *reinterpret_cast<const double2*>(U + c * u_bs + e);
    const double2 v = *reinterpret_cast<const double2*>(V + c * u_bs + e);
    const double a = cs[j], b = sn[j];
    *reinterpret_cast<double2*>(dst + (long long)dst_idx[j] * dst_bs + e) = make_double2(u.x * a + v.x * b, u.y * a + v.y * b);
}

// dst[dst_idx[j]] = src[src_idx[j]]: gathers the chains of a FULL call into the engine's U^T workspace / keeps an accepted proposal
__global__ void __launch_bounds__(256) k_sampler_copy(double* __restrict__ dst, long long dst_bs, const int* __restrict__ dst_idx,
                                                      const double* __restrict__ src, long long src_bs, const int* __restrict__ src_idx,
                                                      long long count) {
    const int j = blockIdx.y;
    const long long e = 2 * ((long long)blockIdx.x * 256 + threadIdx.x);
    if (e >= count) return;
    *reinterpret_cast<double2*>(dst + (long long)dst_idx[j] * dst_bs + e) =
        *reinterpret_cast<const double2*>(src + (long long)src_idx[j] * src_bs + e);
}

}  // namespace apm

// ---- host side ----------------------------------------------------------------------------------------------------
enum { APM_METHOD_MI_MH = 0, APM_METHOD_ESS_MH = 1, APM_METHOD_MI_RDSS = 2, APM_METHOD_ESS_RDSS = 3, APM_METHOD_PMMH = 4 };

namespace {

// FULL_NEW / FULL_GATHER exist once per FULL job slot (+ 2 w): a slot's lists are rewritten only after its previous job finished
enum { IDX_MI = 0, IDX_V, IDX_ELL, IDX_ACC, IDX_FULL_NEW, IDX_FULL_GATHER, IDX_SETS = 8 };
enum { REQ_NONE = 0, REQ_FULL, REQ_FULL_NEWU, REQ_CACHED_NEW, REQ_CACHED_ELL };
enum { ST_START = 0, ST_U_MI, ST_U_ESS, ST_TH_MH, ST_TH_RDSS, ST_PMMH, ST_DONE };
const double S_TWO_PI = 2.0 * 3.14159265358979323846;

struct SChain {
    std::vector<double> theta, theta_req, base, dir;
    double log_f = 0.0;
    int cur = 0, prop = 0;
    int s = 0, state = ST_START, req = REQ_NONE;
    uint64_t seed = 0, n_scalar = 0;
    unsigned n_bulk = 0, draw = 0;     // draw: bulk draw id of the pending request (fresh u / proposal / v)
    double log_y = 0.0, phi = 0.0, lo = 0.0, hi = 0.0, x = 0.0;
    int it = 0;
    bool new_v = false, accept_u = false;
    int64_t n_reject[2] = {0, 0}, n_cubic_ops = 0, n_full = 0, n_cached = 0;
    int failed = 0;
};

struct FullJob {
    std::vector<int> chains, slots, ops, st;
    std::vector<double> thetas, vals;
    int rc = APM_OK;
    std::string err;
};

}  // namespace

struct apm_sampler {
    apm_ctx* eng = nullptr;
    apm_ctx* comp = nullptr;
    apm_ctx* eng2 = nullptr;           // second FULL job: companion with full workspaces on the same cache slots (n_jobs == 2)
    int n_jobs = 1;
    int min_second = 0;                // a second job is only started with at least this many chains
    int method = 0, B = 0, N = 0, Npad = 0, P = 0, n = 0, np = 0, device = 0;
    long long ubs = 0;
    std::vector<uint64_t> seeds;
    std::vector<double> prior_a, prior_b, prior_c, prop_scales;
    double slice_width = 1.0, batch_frac = 0.5;
    int max_slice_iters = 1000;
    double *dU = nullptr, *dV = nullptr, *dCS = nullptr;
    unsigned long long* dSeeds = nullptr;
    // index lists of the device kernels, staged through pinned memory: [IDX_SETS][4 lists][B].  Every set has one user
    // and is rewritten on the host only after a synchronisation of s_main that follows its last upload.
    int* dIdx = nullptr;
    int* hIdx = nullptr;
    double* hCS = nullptr;
    cudaStream_t s_main = nullptr, s_full[2] = {nullptr, nullptr};
    cudaEvent_t ev_full_ready[2] = {nullptr, nullptr};
    std::vector<SChain> chains;
    double* trace = nullptr;
    int n_sample = 0;
    // FULL worker
    std::thread worker[2];
    std::mutex mu;
    std::condition_variable cv;
    FullJob job[2];
    bool job_posted[2] = {false, false}, job_done[2] = {false, false}, quit = false;
    // scheduling diagnostics of the last run
    double stats[8] = {0};
};

namespace {

inline int* s_hidx(apm_sampler* s, int set, int list) { return s->hIdx + ((size_t)set * 4 + list) * s->B; }
inline int* s_didx(apm_sampler* s, int set, int list) { return s->dIdx + ((size_t)set * 4 + list) * s->B; }

// scalar stream of a chain: uniform = u1, normal = sqrt(-2 log u1) cos(2 pi u2) of counter (k, 0, 0, PHILOX_SCALAR)
inline void s_scalar(SChain& ch, double* uni, double* nrm) {
    const apm::Philox4 r = apm::philox4x32_10((uint32_t)ch.n_scalar, (uint32_t)(ch.n_scalar >> 32), 0u, apm::PHILOX_SCALAR,
                                              (uint32_t)ch.seed, (uint32_t)(ch.seed >> 32));
    ch.n_scalar++;
    const double u1 = apm::philox_u01(r.v[0], r.v[1]), u2 = apm::philox_u01(r.v[2], r.v[3]);
    if (uni) *uni = u1;
    if (nrm) *nrm = std::sqrt(-2.0 * std::log(u1)) * std::cos(S_TWO_PI * u2);
}
inline double s_uniform(SChain& ch) { double u; s_scalar(ch, &u, nullptr); return u; }
inline double s_normal(SChain& ch) { double z; s_scalar(ch, nullptr, &z); return z; }

// log prior of theta: sum_k log-Gamma(theta_k; a_k, b_k), gpdemo/utils.py:39-59 as the notebooks' closure adds it (nb cell 12)
inline double s_log_prior(const apm_sampler* s, const double* th) {
    double v = 0.0;
    for (int k = 0; k < s->P; k++) {
        const double t = (s->prior_c[k] + s->prior_a[k] * th[k]) - s->prior_b[k] * std::exp(th[k]);
        v = (k == 0) ? t : v + t;
    }
    return v;
}

void s_fail(SChain& ch, int status) {
    ch.failed = status;
    ch.req = REQ_NONE;
    ch.state = ST_DONE;
}

void s_begin_u(apm_sampler* s, SChain& ch);

void s_finish_iteration(apm_sampler* s, SChain& ch) {
    memcpy(s->trace + ((size_t)(&ch - s->chains.data()) * s->n_sample + ch.s) * s->P, ch.theta.data(), sizeof(double) * s->P);
    ch.s++;
}

void s_begin_pmmh(apm_sampler* s, SChain& ch) {
    if (ch.s >= s->n_sample) { ch.req = REQ_NONE; ch.state = ST_DONE; return; }
    for (int k = 0; k < s->P; k++) ch.theta_req[k] = ch.theta[k] + s->prop_scales[k] * s_normal(ch);   // nb cell 12 prop_sampler
    ch.draw = ch.n_bulk++;                                    // fresh normals inside the estimate closure (smp.py:159-262)
    ch.req = REQ_FULL_NEWU;
    ch.state = ST_PMMH;
}

void s_begin_theta(apm_sampler* s, SChain& ch) {
    const bool mh = s->method == APM_METHOD_MI_MH || s->method == APM_METHOD_ESS_MH;
    if (mh) {
        for (int k = 0; k < s->P; k++) ch.theta_req[k] = ch.theta[k] + s->prop_scales[k] * s_normal(ch);
        ch.req = REQ_FULL;
        ch.state = ST_TH_MH;
        return;
    }
    double nrm = 0.0;                                          // nb-rdss cell 12 dir_and_w_sampler
    for (int k = 0; k < s->P; k++) { ch.dir[k] = s_normal(ch); nrm += ch.dir[k] * ch.dir[k]; }
    nrm = std::sqrt(nrm);
    for (int k = 0; k < s->P; k++) ch.dir[k] /= nrm;
    ch.log_y = std::log(s_uniform(ch)) + ch.log_f;            // mu.py:481
    ch.lo = 0.0 - s->slice_width * s_uniform(ch);             // mu.py:483-484
    ch.hi = ch.lo + s->slice_width;
    ch.base = ch.theta;
    ch.it = 0;
    ch.x = ch.lo + (ch.hi - ch.lo) * s_uniform(ch);           // mu.py:503
    for (int k = 0; k < s->P; k++) ch.theta_req[k] = ch.base[k] + ch.x * ch.dir[k];
    ch.req = REQ_FULL;
    ch.state = ST_TH_RDSS;
}

void s_begin_u(apm_sampler* s, SChain& ch) {
    if (ch.s >= s->n_sample) { ch.req = REQ_NONE; ch.state = ST_DONE; return; }
    const bool mi = s->method == APM_METHOD_MI_MH || s->method == APM_METHOD_MI_RDSS;
    if (mi) {
        ch.draw = ch.n_bulk++;                                 // mu.py:284-288: independent proposal
        ch.req = REQ_CACHED_NEW;
        ch.state = ST_U_MI;
        return;
    }
    ch.draw = ch.n_bulk++;                                     // smp.py:786: v
    ch.log_y = ch.log_f + std::log(s_uniform(ch));            // mu.py:373
    ch.phi = s_uniform(ch) * S_TWO_PI;                        // mu.py:375
    ch.lo = ch.phi - S_TWO_PI;
    ch.hi = ch.phi;
    ch.new_v = true;
    ch.it = 0;
    ch.req = REQ_CACHED_ELL;
    ch.state = ST_U_ESS;
}

// resume a chain with the value of its pending request (estimate + log prior) -- the generators of apm_b200.batched
void s_resume(apm_sampler* s, SChain& ch, double val) {
    switch (ch.state) {
    case ST_START:
        ch.log_f = val;
        ch.s = 1;
        if (s->method == APM_METHOD_PMMH) { s_begin_pmmh(s, ch); return; }
        std::swap(ch.cur, ch.prop);                            // the first estimate's cache is current
        s_begin_u(s, ch);
        return;
    case ST_U_MI:
        if (s_uniform(ch) < std::exp(val - ch.log_f)) {        // mu.py:299-303
            ch.log_f = val;
            ch.accept_u = true;
        } else {
            ch.n_reject[0]++;
        }
        s_begin_theta(s, ch);
        return;
    case ST_U_ESS:
        ch.new_v = false;
        if (val > ch.log_y) {
            ch.log_f = val;
            ch.accept_u = true;
            s_begin_theta(s, ch);
            return;
        }
        if (ch.phi < 0) ch.lo = ch.phi;
        else if (ch.phi > 0) ch.hi = ch.phi;
        else { s_begin_theta(s, ch); return; }                 // slice collapsed (mu.py:391-393)
        if (++ch.it >= s->max_slice_iters) { s_fail(ch, APM_CHAIN_NEWTON_MAXIT); return; }
        ch.phi = ch.lo + s_uniform(ch) * (ch.hi - ch.lo);
        ch.req = REQ_CACHED_ELL;
        return;
    case ST_TH_MH: {
        // symmetric Gaussian random walk: forward and backward proposal densities are the same number (mu.py:149-152)
        double q = 0.0;
        for (int k = 0; k < s->P; k++) { const double d = (ch.theta_req[k] - ch.theta[k]) / s->prop_scales[k]; q += d * d; }
        q *= -0.5;
        if (s_uniform(ch) < std::exp(val + q - ch.log_f - q)) {
            ch.theta = ch.theta_req;
            ch.log_f = val;
            std::swap(ch.cur, ch.prop);                        // smp.py:413-417
        } else {
            ch.n_reject[1]++;
        }
        s_finish_iteration(s, ch);
        s_begin_u(s, ch);
        return;
    }
    case ST_TH_RDSS:
        std::swap(ch.cur, ch.prop);                            // every evaluation overwrites the current cache (smp.py:1083-1085)
        if (val > ch.log_y) {
            ch.theta = ch.theta_req;
            ch.log_f = val;
        } else {
            if (ch.x < 0.) ch.lo = ch.x;
            else if (ch.x > 0.) ch.hi = ch.x;
            else { s_finish_iteration(s, ch); s_begin_u(s, ch); return; }
            if (++ch.it >= s->max_slice_iters) { s_fail(ch, APM_CHAIN_NEWTON_MAXIT); return; }
            ch.x = ch.lo + (ch.hi - ch.lo) * s_uniform(ch);
            for (int k = 0; k < s->P; k++) ch.theta_req[k] = ch.base[k] + ch.x * ch.dir[k];
            ch.req = REQ_FULL;
            return;
        }
        s_finish_iteration(s, ch);
        s_begin_u(s, ch);
        return;
    case ST_PMMH:
    {
        double q = 0.0;
        for (int k = 0; k < s->P; k++) { const double d = (ch.theta_req[k] - ch.theta[k]) / s->prop_scales[k]; q += d * d; }
        q *= -0.5;
        if (s_uniform(ch) < std::exp(val + q - ch.log_f - q)) {   // symmetric random walk (smp.py:223-262)
            ch.theta = ch.theta_req;
            ch.log_f = val;
        } else {
            ch.n_reject[1]++;
        }
        s_finish_iteration(s, ch);
        s_begin_pmmh(s, ch);
        return;
    }
    default:
        ch.req = REQ_NONE;
    }
}

inline apm_ctx* s_engine(apm_sampler* s, int w) { return w == 0 ? s->eng : s->eng2; }

void s_worker(apm_sampler* s, int w) {
    cudaSetDevice(s->device);
    for (;;) {
        std::unique_lock<std::mutex> lk(s->mu);
        s->cv.wait(lk, [s, w] { return s->job_posted[w] || s->quit; });
        if (s->quit) return;
        s->job_posted[w] = false;
        lk.unlock();
        FullJob& j = s->job[w];
        const int m = (int)j.chains.size();
        int rc = (cudaStreamWaitEvent(s->s_full[w], s->ev_full_ready[w], 0) == cudaSuccess) ? APM_OK : APM_ERR_CUDA;
        if (rc == APM_OK)
            rc = estimate_full_impl(s_engine(s, w), j.thetas.data(), nullptr, 1, s->N, m, j.slots.data(), j.vals.data(), j.ops.data(),
                                    j.st.data(), true);
        j.rc = rc;
        if (rc != APM_OK) j.err = g_err;
        lk.lock();
        s->job_done[w] = true;
        lk.unlock();
        s->cv.notify_all();
    }
}

#define S_CU(expr)                                                                \
    do {                                                                          \
        cudaError_t e__ = (expr);                                                 \
        if (e__ != cudaSuccess) {                                                 \
            set_err(std::string(#expr) + ": " + cudaGetErrorString(e__));         \
            return APM_ERR_CUDA;                                                  \
        }                                                                         \
    } while (0)

int s_launch_ok(apm_ctx* c, const char* what) {
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_err(std::string("launch ") + what + ": " + cudaGetErrorString(e));
        return APM_ERR_CUDA;
    }
    return APM_OK;
}

// fresh normals for `cnt` jobs: lists (dst index, chain, draw) in set `set`, lists 0..2, already filled on the host
int s_normals(apm_sampler* s, int set, int cnt, double* dst, long long dst_bs) {
    if (cnt <= 0) return APM_OK;
    S_CU(cudaMemcpyAsync(s_didx(s, set, 0), s_hidx(s, set, 0), sizeof(int) * 3 * (size_t)s->B, cudaMemcpyHostToDevice, s->s_main));
    dim3 grid((s->np + 127) / 128, s->Npad / 2, cnt);
    apm::k_sampler_normals<<<grid, 128, 0, s->s_main>>>(dst, dst_bs, s_didx(s, set, 0), s->dSeeds, s_didx(s, set, 1),
                                                        reinterpret_cast<const unsigned*>(s_didx(s, set, 2)), s->n, s->N, s->np);
    return s_launch_ok(s->comp, "k_sampler_normals");
}

int s_copy(apm_sampler* s, int set, int cnt, double* dst, const double* src) {
    if (cnt <= 0) return APM_OK;
    S_CU(cudaMemcpyAsync(s_didx(s, set, 0), s_hidx(s, set, 0), sizeof(int) * 2 * (size_t)s->B, cudaMemcpyHostToDevice, s->s_main));
    dim3 grid((unsigned)((s->ubs / 2 + 255) / 256), cnt);
    apm::k_sampler_copy<<<grid, 256, 0, s->s_main>>>(dst, s->ubs, s_didx(s, set, 0), src, s->ubs, s_didx(s, set, 1), s->ubs);
    return s_launch_ok(s->comp, "k_sampler_copy");
}

// gather the chains of the next FULL call (fresh normals first where the request asks for them) and hand it to the worker
int s_submit_full(apm_sampler* s, const std::vector<int>& full, int w) {
    const int m = (int)full.size();
    FullJob& j = s->job[w];
    j.chains = full;
    j.thetas.resize((size_t)m * s->P); j.slots.resize(m); j.vals.resize(m); j.ops.resize(m); j.st.resize(m);
    int n_new = 0;
    for (int q = 0; q < m; q++) {
        SChain& ch = s->chains[full[q]];
        if (ch.req == REQ_FULL_NEWU) {
            s_hidx(s, IDX_FULL_NEW + 2 * w, 0)[n_new] = full[q];      // fresh u of the chain (PM-MH: every estimate; APM: the first)
            s_hidx(s, IDX_FULL_NEW + 2 * w, 1)[n_new] = full[q];
            s_hidx(s, IDX_FULL_NEW + 2 * w, 2)[n_new] = (int)ch.draw;
            n_new++;
        }
        memcpy(&j.thetas[(size_t)q * s->P], ch.theta_req.data(), sizeof(double) * s->P);
        j.slots[q] = ch.prop;                                   // written into the proposal slot
    }
    APM_TRY(s_normals(s, IDX_FULL_NEW + 2 * w, n_new, s->dU, s->ubs));
    for (int q = 0; q < m; q++) {
        s_hidx(s, IDX_FULL_GATHER + 2 * w, 0)[q] = q;
        s_hidx(s, IDX_FULL_GATHER + 2 * w, 1)[q] = full[q];
    }
    APM_TRY(s_copy(s, IDX_FULL_GATHER + 2 * w, m, s_engine(s, w)->dUT, s->dU));
    S_CU(cudaEventRecord(s->ev_full_ready[w], s->s_main));
    {
        std::lock_guard<std::mutex> lk(s->mu);
        s->job_done[w] = false;
        s->job_posted[w] = true;
    }
    s->cv.notify_all();
    return APM_OK;
}

int s_run(apm_sampler* s, const double* theta_init, int n_sample, double* thetas_out, int64_t* counts_out) {
    const int B = s->B, P = s->P;
    S_CU(cudaSetDevice(s->device));
    s->n_sample = n_sample;
    s->trace = thetas_out;
    for (size_t e = 0; e < (size_t)B * n_sample * P; e++) thetas_out[e] = NAN;
    s->chains.assign(B, SChain());
    int n_pending = 0;
    for (int c = 0; c < B; c++) {
        SChain& ch = s->chains[c];
        ch.theta.assign(theta_init + (size_t)c * P, theta_init + (size_t)(c + 1) * P);
        ch.theta_req = ch.theta; ch.base = ch.theta; ch.dir.assign(P, 0.0);
        ch.seed = s->seeds[c];
        ch.cur = 2 * c; ch.prop = 2 * c + 1;
        memcpy(thetas_out + (size_t)c * n_sample * P, ch.theta.data(), sizeof(double) * P);     // trace[0]
        ch.draw = ch.n_bulk++;                                   // smp.py:377 / 546 / 689 / 825: the initial u
        ch.req = REQ_FULL_NEWU;
        ch.state = ST_START;
        n_pending++;
    }
    // the engines run the FULL calls on their workers' streams, the companion the CACHED calls on this thread's
    const int J = s->n_jobs;
    const cudaStream_t eng_stream = s->eng->stream;
    S_CU(cudaStreamSynchronize(eng_stream));
    s->eng->stream = s->s_full[0];
    s->comp->stream = s->s_main;
    if (s->eng2) {      // the second engine follows the settings of the first
        apm_ctx *a = s->eng, *b2 = s->eng2;
        b2->stream = s->s_full[1];
        b2->tol = a->tol; b2->max_iters = a->max_iters; b2->approx = a->approx;
        b2->ep_tol = a->ep_tol; b2->ep_max_iters = a->ep_max_iters; b2->ep_damping = a->ep_damping;
        b2->overlap_chol_k = a->overlap_chol_k;
    }
    bool inflight[2] = {false, false};
    struct Restore {     // also on error returns: wait for the FULL calls still in flight before handing the engine back
        apm_sampler* s; cudaStream_t st; bool* inflight;
        ~Restore() {
            for (int w = 0; w < 2; w++) {
                if (inflight[w]) { std::unique_lock<std::mutex> lk(s->mu); s->cv.wait(lk, [this, w] { return s->job_done[w]; }); }
                if (s->s_full[w]) cudaStreamSynchronize(s->s_full[w]);
            }
            cudaStreamSynchronize(s->s_main);
            s->eng->stream = st;
        }
    } restore{s, eng_stream, inflight};
    const auto t_start = std::chrono::steady_clock::now();
    auto now_s = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count(); };
    double t_submit[2] = {0.0, 0.0}, t_flight = 0.0, t_busy = 0.0, t_busy_from = 0.0;
    int64_t full_calls = 0, full_chains = 0, cached_calls = 0, cached_chains = 0, rounds = 0;
    std::vector<int> full, cached, slots, st;
    std::vector<double> vals;
    auto count_pending = [&] {
        int k = 0;
        for (const SChain& ch : s->chains) k += ch.req != REQ_NONE;
        return k;
    };
    auto any_inflight = [&] { return inflight[0] || inflight[1]; };
    while (n_pending > 0 || any_inflight()) {
        rounds++;
        int n_cached_req = 0, n_full_req = 0;
        for (const SChain& ch : s->chains) {
            n_cached_req += (ch.req == REQ_CACHED_NEW || ch.req == REQ_CACHED_ELL);
            n_full_req += (ch.req == REQ_FULL || ch.req == REQ_FULL_NEWU);
        }
        // ---- finished FULL calls are harvested as soon as they are seen; their chains are resumed at once so that those
        // which need another FULL estimate join the next call.  With nothing else to do (no CACHED work, and no FULL call
        // that could be submitted now) the thread sleeps until a call finishes.
        if (any_inflight()) {
            const bool free_slot = (J > 1) && !(inflight[0] && inflight[1]);
            const bool can_submit = free_slot && n_full_req >= std::max(1, s->min_second) && n_full_req >= s->batch_frac * n_pending;
            std::unique_lock<std::mutex> lk(s->mu);
            if (n_cached_req == 0 && !can_submit)
                s->cv.wait(lk, [&] { return (inflight[0] && s->job_done[0]) || (inflight[1] && s->job_done[1]); });
            bool done[2] = {inflight[0] && s->job_done[0], inflight[1] && s->job_done[1]};
            lk.unlock();
            for (int w = 0; w < 2; w++) {
                if (!done[w]) continue;
                inflight[w] = false;
                t_flight += now_s() - t_submit[w];
                if (!any_inflight()) t_busy += now_s() - t_busy_from;
                FullJob& j = s->job[w];
                if (j.rc != APM_OK) { set_err("FULL call: " + j.err); return j.rc; }
                for (size_t q = 0; q < j.chains.size(); q++) {
                    SChain& ch = s->chains[j.chains[q]];
                    ch.n_full++;
                    if (j.st[q] != 0) { s_fail(ch, j.st[q]); continue; }
                    ch.n_cubic_ops += j.ops[q];
                    s_resume(s, ch, j.vals[q] + s_log_prior(s, &j.thetas[q * (size_t)P]));
                }
            }
        }
        // ---- submit the next FULL call to a free job slot: enough chains are waiting for one, or nobody has CACHED work left.
        // A second call beside one in flight is only worth its fixed cost with at least min_second chains.
        {
            int w = -1;
            for (int q = 0; q < J; q++)
                if (!inflight[q]) { w = q; break; }
            if (w >= 0) {
                full.clear();
                int n_pend = 0;
                for (int c = 0; c < B; c++) {
                    const int r = s->chains[c].req;
                    n_pend += r != REQ_NONE;
                    if (r == REQ_FULL || r == REQ_FULL_NEWU) full.push_back(c);
                }
                const bool alone = !any_inflight();
                const bool enough = (double)full.size() >= s->batch_frac * n_pend || (int)full.size() == n_pend;
                if (!full.empty() && enough && (alone || (int)full.size() >= s->min_second)) {
                    APM_TRY(s_submit_full(s, full, w));
                    for (int c : full) s->chains[c].req = REQ_NONE;     // in flight: not pending
                    if (alone) t_busy_from = now_s();
                    inflight[w] = true;
                    t_submit[w] = now_s();
                    full_calls++;
                    full_chains += (int64_t)full.size();
                }
            }
        }
        // ---- CACHED requests of the chains that are not in flight
        cached.clear();
        for (int c = 0; c < B; c++)
            if (s->chains[c].req == REQ_CACHED_NEW || s->chains[c].req == REQ_CACHED_ELL) cached.push_back(c);
        if (!cached.empty()) {
            const int m = (int)cached.size();
            // independent proposals straight into the companion's U^T workspace; fresh v of the ellipse into dV
            int n_mi = 0, n_v = 0, n_ell = 0;
            for (int q = 0; q < m; q++) {
                SChain& ch = s->chains[cached[q]];
                if (ch.req == REQ_CACHED_NEW) {
                    s_hidx(s, IDX_MI, 0)[n_mi] = q; s_hidx(s, IDX_MI, 1)[n_mi] = cached[q]; s_hidx(s, IDX_MI, 2)[n_mi] = (int)ch.draw;
                    n_mi++;
                } else if (ch.new_v) {
                    s_hidx(s, IDX_V, 0)[n_v] = cached[q]; s_hidx(s, IDX_V, 1)[n_v] = cached[q]; s_hidx(s, IDX_V, 2)[n_v] = (int)ch.draw;
                    n_v++;
                }
            }
            APM_TRY(s_normals(s, IDX_MI, n_mi, s->comp->dUT, s->ubs));
            APM_TRY(s_normals(s, IDX_V, n_v, s->dV, s->ubs));
            for (int q = 0; q < m; q++) {
                SChain& ch = s->chains[cached[q]];
                if (ch.req != REQ_CACHED_ELL) continue;
                s_hidx(s, IDX_ELL, 0)[n_ell] = q;
                s_hidx(s, IDX_ELL, 1)[n_ell] = cached[q];
                s->hCS[n_ell] = std::cos(ch.phi);
                s->hCS[B + n_ell] = std::sin(ch.phi);
                n_ell++;
            }
            if (n_ell) {
                S_CU(cudaMemcpyAsync(s_didx(s, IDX_ELL, 0), s_hidx(s, IDX_ELL, 0), sizeof(int) * 2 * B, cudaMemcpyHostToDevice, s->s_main));
                S_CU(cudaMemcpyAsync(s->dCS, s->hCS, sizeof(double) * 2 * B, cudaMemcpyHostToDevice, s->s_main));
                dim3 grid((unsigned)((s->ubs / 2 + 255) / 256), n_ell);
                apm::k_sampler_ellipse<<<grid, 256, 0, s->s_main>>>(s->comp->dUT, s->ubs, s_didx(s, IDX_ELL, 0), s->dU, s->dV, s->ubs,
                                                                   s_didx(s, IDX_ELL, 1), s->dCS, s->dCS + B, s->ubs);
                APM_TRY(s_launch_ok(s->comp, "k_sampler_ellipse"));
            }
            slots.resize(m); st.resize(m); vals.resize(m);
            for (int q = 0; q < m; q++) slots[q] = s->chains[cached[q]].cur;
            APM_TRY(cached_common(s->comp, slots.data(), nullptr, 1, s->N, m, vals.data(), nullptr, st.data(), true));
            cached_calls++;
            cached_chains += m;
            int n_acc = 0;
            for (int q = 0; q < m; q++) {
                SChain& ch = s->chains[cached[q]];
                ch.n_cached++;
                if (st[q] != 0) { s_fail(ch, st[q]); continue; }
                s_resume(s, ch, vals[q] + s_log_prior(s, ch.theta.data()));
                if (ch.accept_u) {
                    ch.accept_u = false;
                    s_hidx(s, IDX_ACC, 0)[n_acc] = cached[q];        // dst: the chain's u
                    s_hidx(s, IDX_ACC, 1)[n_acc] = q;                // src: its proposal in the companion's workspace
                    n_acc++;
                }
            }
            APM_TRY(s_copy(s, IDX_ACC, n_acc, s->dU, s->comp->dUT));
        }
        n_pending = count_pending();
    }
    S_CU(cudaStreamSynchronize(s->s_main));
    for (int c = 0; c < B; c++) {
        const SChain& ch = s->chains[c];
        int64_t* o = counts_out + (size_t)c * 6;
        o[0] = ch.n_reject[0]; o[1] = ch.n_reject[1]; o[2] = ch.n_cubic_ops; o[3] = ch.n_full; o[4] = ch.n_cached; o[5] = ch.failed;
    }
    s->stats[0] = (double)full_calls; s->stats[1] = (double)full_chains; s->stats[2] = (double)cached_calls;
    s->stats[3] = (double)cached_chains; s->stats[4] = t_flight; s->stats[5] = now_s(); s->stats[6] = (double)rounds;
    s->stats[7] = t_busy;      // time with at least one FULL call in flight (stats[4] sums the calls' own durations)
    return APM_OK;
}

}  // namespace

extern "C" int apm_sampler_destroy(apm_sampler* s) {
    if (!s) return APM_OK;
    cudaSetDevice(s->device);
    {
        std::lock_guard<std::mutex> lk(s->mu);
        s->quit = true;
    }
    s->cv.notify_all();
    for (int w = 0; w < 2; w++)
        if (s->worker[w].joinable()) s->worker[w].join();
    if (s->s_main) cudaStreamSynchronize(s->s_main);
    for (int w = 0; w < 2; w++)
        if (s->s_full[w]) cudaStreamSynchronize(s->s_full[w]);
    if (s->comp) apm_destroy(s->comp);
    if (s->eng2) apm_destroy(s->eng2);
    if (s->dU) cudaFree(s->dU);
    if (s->dV) cudaFree(s->dV);
    if (s->dCS) cudaFree(s->dCS);
    if (s->dSeeds) cudaFree(s->dSeeds);
    if (s->dIdx) cudaFree(s->dIdx);
    if (s->hIdx) cudaFreeHost(s->hIdx);
    if (s->hCS) cudaFreeHost(s->hCS);
    for (int w = 0; w < 2; w++) {
        if (s->ev_full_ready[w]) cudaEventDestroy(s->ev_full_ready[w]);
        if (s->s_full[w]) cudaStreamDestroy(s->s_full[w]);
    }
    if (s->s_main) cudaStreamDestroy(s->s_main);
    delete s;
    return APM_OK;
}

extern "C" int apm_sampler_create(apm_ctx* ctx, int method, int n_chains, int n_imp, const uint64_t* seeds, const double* prior_ab,
                                  const double* prop_scales, double slice_width, int max_slice_iters, apm_sampler** out) {
    if (!ctx || !seeds || !prior_ab || !out || n_chains <= 0 || method < APM_METHOD_MI_MH || method > APM_METHOD_PMMH) {
        set_err("apm_sampler_create: invalid argument");
        return APM_ERR_INVALID;
    }
    APM_TRY(not_companion(ctx));
    const bool mh = method == APM_METHOD_MI_MH || method == APM_METHOD_ESS_MH || method == APM_METHOD_PMMH;
    if (mh && !prop_scales) {
        set_err("apm_sampler_create: the Metropolis-Hastings theta-updates need prop_scales");
        return APM_ERR_INVALID;
    }
    if (n_chains > ctx->maxB || 2 * n_chains > ctx->nslots || n_imp <= 0 || n_imp > ctx->maxN) {
        set_err("apm_sampler_create: the engine context needs max_chains >= n_chains, n_slots >= 2 n_chains and max_nimp >= n_imp");
        return APM_ERR_INVALID;
    }
    S_CU(cudaSetDevice(ctx->device));
    apm_sampler* s = new apm_sampler();
    s->eng = ctx; s->device = ctx->device;
    s->method = method; s->B = n_chains; s->N = n_imp; s->Npad = (n_imp + TB - 1) / TB * TB;
    s->P = ctx->P; s->n = ctx->n; s->np = ctx->np;
    s->ubs = (long long)s->Npad * s->np;
    s->seeds.assign(seeds, seeds + n_chains);
    for (int k = 0; k < s->P; k++) {
        const double a = prior_ab[2 * k], b = prior_ab[2 * k + 1];
        s->prior_a.push_back(a); s->prior_b.push_back(b);
        s->prior_c.push_back(a * std::log(b) - std::lgamma(a));
    }
    if (prop_scales) s->prop_scales.assign(prop_scales, prop_scales + s->P);
    s->slice_width = slice_width;
    s->max_slice_iters = max_slice_iters > 0 ? max_slice_iters : 1000;
    if (getenv("APM_SAMPLER_BATCH_FRAC") && atof(getenv("APM_SAMPLER_BATCH_FRAC")) > 0) s->batch_frac = atof(getenv("APM_SAMPLER_BATCH_FRAC"));
    // FULL calls in flight at once (1 or 2) and the smallest call worth starting beside another one
    // (default 1: two persistent factorisations in flight do not share the SMs' CTA slots elastically -- measured +-2 %)
    s->n_jobs = (getenv("APM_SAMPLER_JOBS") && atoi(getenv("APM_SAMPLER_JOBS")) == 2) ? 2 : 1;
    s->min_second = std::max(8, n_chains / 5);
    if (getenv("APM_SAMPLER_MIN_SECOND") && atoi(getenv("APM_SAMPLER_MIN_SECOND")) > 0) s->min_second = atoi(getenv("APM_SAMPLER_MIN_SECOND"));
    const size_t B = n_chains;
    const bool ess = method == APM_METHOD_ESS_MH || method == APM_METHOD_ESS_RDSS;
    int prio_lo = 0, prio_hi = 0;      // numerically lower = higher priority
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (getenv("APM_SAMPLER_NO_PRIORITY")) prio_hi = prio_lo = 0;
    bool ok = cudaMalloc(&s->dU, sizeof(double) * B * s->ubs) == cudaSuccess &&
              (!ess || cudaMalloc(&s->dV, sizeof(double) * B * s->ubs) == cudaSuccess) &&
              cudaMalloc(&s->dCS, sizeof(double) * 2 * B) == cudaSuccess &&
              cudaMalloc(&s->dSeeds, sizeof(unsigned long long) * B) == cudaSuccess &&
              cudaMalloc(&s->dIdx, sizeof(int) * IDX_SETS * 4 * B) == cudaSuccess &&
              cudaMallocHost(&s->hIdx, sizeof(int) * IDX_SETS * 4 * B) == cudaSuccess &&
              cudaMallocHost(&s->hCS, sizeof(double) * 2 * B) == cudaSuccess &&
              // the FULL calls bound the run, the CACHED calls only have to be back before the next FULL call starts: blocks of
              // the FULL streams are scheduled first whenever both have work pending
              cudaStreamCreateWithPriority(&s->s_main, cudaStreamNonBlocking, prio_lo) == cudaSuccess &&
              cudaStreamCreateWithPriority(&s->s_full[0], cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
              cudaStreamCreateWithPriority(&s->s_full[1], cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
              cudaEventCreateWithFlags(&s->ev_full_ready[0], cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&s->ev_full_ready[1], cudaEventDisableTiming) == cudaSuccess &&
              cudaMemcpy(s->dSeeds, seeds, sizeof(unsigned long long) * B, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        set_err(std::string("apm_sampler_create: allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
        apm_sampler_destroy(s);
        return APM_ERR_NOMEM;
    }
    int rc = apm_create_companion(ctx, n_chains, n_imp, &s->comp);
    if (rc != APM_OK) {
        apm_sampler_destroy(s);
        return rc;
    }
    if (s->n_jobs > 1) {
        rc = create_full_companion(ctx, n_chains, n_imp, &s->eng2);
        if (rc != APM_OK) {
            apm_sampler_destroy(s);
            return rc;
        }
    }
    // the engines' auxiliary streams (chol K, M-space form of a mixed Newton round) belong to the FULL calls: same priority
    for (apm_ctx* e : {s->eng, s->eng2}) {
        if (!e || prio_hi == prio_lo) continue;
        cudaStream_t hi = nullptr;
        if (cudaStreamCreateWithPriority(&hi, cudaStreamNonBlocking, prio_hi) == cudaSuccess) {
            cudaStreamSynchronize(e->aux_stream);
            cudaStreamDestroy(e->aux_stream);
            e->aux_stream = hi;
        }
    }
    for (int w = 0; w < s->n_jobs; w++) s->worker[w] = std::thread(s_worker, s, w);
    *out = s;
    return APM_OK;
}

extern "C" int apm_sampler_run(apm_sampler* s, const double* theta_init, int n_sample, double* thetas_out, int64_t* counts_out) {
    if (!s || !theta_init || !thetas_out || !counts_out || n_sample <= 0) {
        set_err("apm_sampler_run: invalid argument");
        return APM_ERR_INVALID;
    }
    return s_run(s, theta_init, n_sample, thetas_out, counts_out);
}

extern "C" int apm_sampler_stats(apm_sampler* s, double* out8) {
    if (!s || !out8) return APM_ERR_INVALID;
    memcpy(out8, s->stats, sizeof(s->stats));
    return APM_OK;
}
