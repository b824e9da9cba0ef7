// apm_capi.cu -- context, orchestration and the C ABI declared in include/apm_b200.h.
//
// One context = one data set on one GPU.  All bulk state lives in HBM:
//   K, LB, Z        [max_chains][np][np]   prior covariance, chol(B) of the current Newton step, Z = (L_B^{-1} W^1/2 K)^T
//   slot L_K, L_C   [n_slots][np][np]      the reference's cached_results (chol K, chol C), row-major lower
//   slot mu         [n_slots][np]          f_post
//   UT, F, Zf       [max_chains][Npad][np] transposed auxiliary normals, latent samples, L_K^{-1} f
// np = n rounded up to 64; padded rows/cols carry the identity so every kernel works on whole tiles.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <string>
#include <vector>

#include "../../include/apm_b200.h"
#include "tile_engine.cuh"
#include "chol_flow.cuh"
#include "tmap_host.h"
#include "vec_kernels.cuh"

using namespace apm;

static thread_local std::string g_err;
static void set_err(const std::string& s) { g_err = s; }

#define CU_TRY(expr)                                                                                  \
    do {                                                                                              \
        cudaError_t e__ = (expr);                                                                     \
        if (e__ != cudaSuccess) {                                                                     \
            set_err(std::string(#expr) + ": " + cudaGetErrorString(e__));                             \
            return (e__ == cudaErrorMemoryAllocation) ? APM_ERR_NOMEM : APM_ERR_CUDA;                 \
        }                                                                                             \
    } while (0)
#define APM_TRY(expr)                 \
    do {                              \
        int r__ = (expr);             \
        if (r__ != APM_OK) return r__; \
    } while (0)

enum { V_F = 0, V_W, V_WS, V_B, V_A, V_T, V_S, V_FNEW, V_S2, V_COUNT };

// kernel ids for launch accounting / profiling
enum { KID_BUILD_K = 0, KID_CHOL, KID_TRSM, KID_SYRK, KID_GEMM_TRI, KID_MATVEC, KID_TRSV, KID_NEWTON_VEC, KID_EPILOGUE,
       KID_TRANSPOSE, KID_MISC, KID_COUNT };
static const char* const KID_NAMES[KID_COUNT] = {"k_build_K", "k_chol", "k_trsm_rows", "k_syrk_sub", "k_gemm_tri",
                                                 "k_matvec", "k_trsv2", "k_newton_vec", "k_is_epilogue",
                                                 "k_transpose_u", "misc"};

struct apm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // H2D of the auxiliary normals overlaps the O(n^3) front of a FULL estimate
    cudaEvent_t copy_done = nullptr;
    bool u_staged = false;
    // host u of the running apm_estimate_full: its upload is started by run_newton right before the first host round trip,
    // i.e. after the whole mode search has been queued (a pageable buffer makes cudaMemcpyAsync block the host thread)
    const double* pend_u = nullptr; int pend_N = 0, pend_B = 0;
    // chol(K) is only needed by the importance-sampling tail: it runs on aux_stream (per-step launches) as filler
    // work beside the Newton rounds, whose latency-bound kernels leave SM slots idle
    cudaStream_t aux_stream = nullptr;
    cudaStream_t launch_stream = nullptr;   // stream the launch helpers (run_chol, profiling events) currently target
    cudaEvent_t ev_k_ready = nullptr, ev_lk_done = nullptr;
    cudaEvent_t ev_mix_fork = nullptr, ev_mix_join = nullptr;   // mixed Newton round: the minority form runs on aux_stream
    bool overlap_chol_k = true;
    bool factored_cov = true;   // chol(C) = L_K U^-T (n^3) instead of TRSM + SYRK + chol (7/3 n^3); APM_EXPLICIT_COV=1 -> reference formulation
    int n = 0, D = 0, np = 0, nb = 0, P = 0, kind = 0;
    double eps = 1e-8;
    int maxB = 0, nslots = 0, maxN = 0, maxNpad = 0;
    double tol = 1e-4;
    int max_iters = 1000;
    // posterior approximation of the FULL estimate: 0 = Laplace (the reference), 1 = EP (extension)
    int approx = 0;
    double ep_tol = 1e-6, ep_damping = 1.0;
    int ep_max_iters = 100;
    double* dEpDelta = nullptr;
    // hybrid Newton: the iteration predicted to be a chain's last is done in M-space (M = I + L_K^T W L_K), whose
    // Cholesky factor is exactly what the factored covariance needs, so the converged chain skips the separate
    // SYRK + Cholesky of M' (and the chol(B) of that iteration).  APM_NO_HYBRID_NEWTON=1 disables it.
    bool hybrid_newton = true;
    int flow_grid_small = 0;            // persistent grid of the 2-CTAs-per-SM instantiations of k_chol_flow
    int flow_small_max = 200;           // batches of at most this many chains use them (APM_FLOW_SMALL_MAX; 0: never)
    int newton_late_round = 4;          // Newton rounds from this index on (0-based) are expected to hold only stragglers (APM_NEWTON_LATE_ROUND)
    int flow_spin_ns = 64;              // back-off of the producer lanes' dependency polls in k_chol_flow (APM_FLOW_SPIN_NS)
    bool fused_fwd = true;              // forward substitution of the Newton solves inside k_chol_flow's diagonal tasks (APM_NO_FUSED_FWD=1: k_trsv2 does both halves)
    int trsv_cluster_max = 160;         // backward solve by a cluster of 4 CTAs per chain for batches of at most this many chains (APM_TRSV_CLUSTER_MAX; 0: never)
    bool fused_vt = true;               // k_chol_flow<true> also stores V = anti-transpose of L' (APM_NO_FUSED_VT=1: separate k_antitranspose)
    double pred_factor = 0.15;   // measured optimum 0.1-0.2 (profiles/): a missed prediction costs a latency-bound covariance phase
    int *dMaskM = nullptr, *dMaskB = nullptr, *dDoneM = nullptr;
    // f_new = s / W^1/2 instead of the second mat-vec of a B-space Newton step (k_fnew_from_s); APM_FNEW_THR=0 disables it
    double fnew_thr = 1e-2;
    bool newton_b_finishers = true;   // set by run_newton: some chain finished in a B-space round (needs the covariance phase)
    size_t mat = 0;  // np*np
    double *dX = nullptr, *dy = nullptr;
    double *dK = nullptr, *dLB = nullptr, *dZ = nullptr;
    double *dSlotLK = nullptr, *dSlotLC = nullptr, *dSlotMu = nullptr, *dSlotLdK = nullptr, *dSlotLdC = nullptr;
    // Factored cache: what a slot's "L_C" buffer holds.  0: chol(C) itself (imported / explicit covariance);
    // 1: V = U^T with M = I + L_K^T W L_K = U U^T, so that L_C = L_K V^-1 is never formed: f_s = mu + L_K (V^-T u_s) and
    // L_K^-1 f_s = L_K^T a + V^-T u_s (mu = K a), with dSlotMt = L_K^T a.  See run_covariance_factored / run_is_tail.
    double* dSlotMt = nullptr;
    std::vector<char> slot_mode;
    double* dLdB = nullptr;
    int* dFlowCounter = nullptr;   // per queue set: [0] task queue head, [1] chains to factorise (k_chol_flow_init)
    // k_chol_flow (TMA / mbarrier dataflow Cholesky): tensor maps over the three matrix buffers it factors into, and two
    // sets of queue state + packed diagonal blocks (set 1: launches on the aux stream, which overlap the main stream's)
    CUtensorMap tmLB, tmSlotLK, tmSlotLC, tmK, tmSlotLK16;   // (16: k-major 16x16 boxes of L_K for the fused M' source)
    bool tma_ok = false;
    int *dFlow2Progress[2] = {nullptr, nullptr}, *dFlow2Skip[2] = {nullptr, nullptr}, *dFlowYProg[2] = {nullptr, nullptr};
    double* dDiagPack[2] = {nullptr, nullptr};
    int flow2_grid = 0;           // resident CTAs of k_chol_flow on the whole GPU
    unsigned long long* dWork = nullptr;   // [0] chain-Choleskys, [1] M' builds executed (counted on the device: the masks are only known there)
    int newton_r0 = 5;            // Newton rounds queued before the first host round trip (APM_NEWTON_R0)
    double *dSymvDirect = nullptr, *dSymvPart = nullptr;   // scratch of the symmetric mat-vec
    double* dInvB = nullptr;   // (L_kk^{-1})^T diagonal blocks of chol(B): [max_chains][nb][64*64]
    double* dVec[V_COUNT] = {nullptr};
    double *dUT = nullptr, *dF = nullptr, *dZf = nullptr, *dUstage = nullptr;
    double *dKp = nullptr, *dOut = nullptr, *dLogw = nullptr;
    int *dStatus = nullptr, *dActive = nullptr, *dIters = nullptr, *dNActive = nullptr, *dSlotsA = nullptr, *dSlotsB = nullptr;
    // pinned host staging
    double *hKp = nullptr, *hOut = nullptr;
    int *hInts = nullptr, *hNActive = nullptr;
    std::vector<char> slot_valid;
    apm_ctx* root = nullptr;            // companion context (apm_create_companion): the slot owner
    bool cached_only = false;           // companion context: shares the parent's cache slots, cached estimates only
    int64_t launches = 0;
    int newton_b_finisher_count = 0;   // set by run_newton: chains that finished in a B-space round
    std::vector<void*> allocs;
    // optional per-kernel CUDA-event timing (apm_profile): events bracket every launch on ctx->stream
    bool prof = false;
    std::vector<cudaEvent_t> ev_pool;
    struct Pending { int kid; cudaEvent_t a, b; };
    std::vector<Pending> pending;
    double prof_ms[KID_COUNT] = {0};
    int64_t prof_n[KID_COUNT] = {0};
};

static char* slot_flags(apm_ctx* c) { return (c->root ? c->root : c)->slot_valid.data(); }
static char* slot_modes(apm_ctx* c) { return (c->root ? c->root : c)->slot_mode.data(); }

template <typename T>
static int dev_alloc(apm_ctx* c, T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T) > 0 ? count * sizeof(T) : 16);
    if (e != cudaSuccess) {
        set_err(std::string("cudaMalloc of ") + std::to_string(count * sizeof(T)) + " bytes: " + cudaGetErrorString(e));
        return APM_ERR_NOMEM;
    }
    c->allocs.push_back(q);
    *p = (T*)q;
    return APM_OK;
}

static cudaEvent_t prof_event(apm_ctx* c) {
    if (!c->ev_pool.empty()) {
        cudaEvent_t e = c->ev_pool.back();
        c->ev_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
static void prof_begin(apm_ctx* c, int kid) {
    if (!c->prof) return;
    apm_ctx::Pending p;
    p.kid = kid;
    p.a = prof_event(c);
    p.b = prof_event(c);
    cudaEventRecord(p.a, c->launch_stream ? c->launch_stream : c->stream);
    c->pending.push_back(p);
}
// fold finished launches into the per-kernel totals (call only after a stream synchronisation)
static void prof_resolve(apm_ctx* c) {
    for (auto& p : c->pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            c->prof_ms[p.kid] += ms;
            c->prof_n[p.kid] += 1;
        } else {
            cudaGetLastError();
        }
        c->ev_pool.push_back(p.a);
        c->ev_pool.push_back(p.b);
    }
    c->pending.clear();
}
static int check_launch(apm_ctx* c, const char* what) {
    c->launches++;
    if (c->prof && !c->pending.empty()) cudaEventRecord(c->pending.back().b, c->launch_stream ? c->launch_stream : c->stream);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_err(std::string("launch ") + what + ": " + cudaGetErrorString(e));
        return APM_ERR_CUDA;
    }
    return APM_OK;
}

static int g_attr_done = 0;
static int set_kernel_attrs() {
    if (g_attr_done) return APM_OK;
    CU_TRY(cudaFuncSetAttribute(k_chol_flow<false, false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF_SMEM_BYTES));
    CU_TRY(cudaFuncSetAttribute(k_chol_flow<true, false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF_SMEM_BYTES));
    CU_TRY(cudaFuncSetAttribute(k_chol_flow<false, true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF_SMEM_BYTES));
    CU_TRY(cudaFuncSetAttribute(k_chol_flow<true, true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF_SMEM_BYTES));
    CU_TRY(cudaFuncSetAttribute(k_chol_flow<false, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF_SMEM_BYTES));
    CU_TRY(cudaFuncSetAttribute(k_chol_flow<true, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF_SMEM_BYTES));
    CU_TRY(cudaFuncSetAttribute(k_chol_flow<false, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF_SMEM_BYTES));
    CU_TRY(cudaFuncSetAttribute(k_chol_flow<true, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF_SMEM_BYTES));
    CU_TRY(cudaFuncSetAttribute(k_trsm_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM_BYTES));
    CU_TRY(cudaFuncSetAttribute(k_syrk_sub, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM_BYTES));
    CU_TRY(cudaFuncSetAttribute(k_trsm_rev, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM_BYTES));
    CU_TRY(cudaFuncSetAttribute(k_gemm_tri, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM_BYTES));
    CU_TRY(cudaFuncSetAttribute(k_build_K, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CU_TRY(cudaFuncSetAttribute(k_build_dK, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CU_TRY(cudaFuncSetAttribute(k_trsv2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CU_TRY(cudaFuncSetAttribute(k_trsv2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CU_TRY(cudaFuncSetAttribute(k_trsv_back_c4, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CU_TRY(cudaFuncSetAttribute(k_is_epilogue, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    g_attr_done = 1;
    return APM_OK;
}

extern "C" const char* apm_version(void) { return "apm_b200 0.1.0 (sm_100a, fp64 DMMA tile engine)"; }
extern "C" const char* apm_last_error(void) { return g_err.c_str(); }

extern "C" int apm_destroy(apm_ctx* c);

// parent != null: companion context -- its cache slots ARE the parent's (no slot storage of its own) and its per-chain
// matrix workspaces are sized for one chain (it only runs the O(n^2 N) cached estimates)
// full_ws (companions only): matrix workspaces for every chain, so that the companion can run FULL estimates into the parent's
// slots as well (second FULL job of the native sampler)
static int create_impl(const double* X, const double* y, int n, int D, int kernel_kind, double epsilon, int max_chains,
                       int n_slots, int max_nimp, int device, apm_ctx* parent, apm_ctx** out, bool full_ws = false) {
    if (!X || !y || !out || n <= 0 || D <= 0 || max_chains <= 0 || n_slots <= 0 || max_nimp <= 0 ||
        (kernel_kind != APM_KERNEL_ISO && kernel_kind != APM_KERNEL_ARD)) {
        set_err("apm_create: invalid argument");
        return APM_ERR_INVALID;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_err("apm_create: no CUDA device visible (apm_b200 has no CPU fallback)");
        return APM_ERR_NOGPU;
    }
    if (device < 0 || device >= ndev) {
        set_err("apm_create: bad device ordinal");
        return APM_ERR_INVALID;
    }
    CU_TRY(cudaSetDevice(device));
    APM_TRY(set_kernel_attrs());
    apm_ctx* c = new apm_ctx();
    c->device = device;
    c->n = n;
    c->D = D;
    c->np = (n + TB - 1) / TB * TB;
    c->nb = c->np / TB;
    c->kind = kernel_kind;
    c->P = (kernel_kind == APM_KERNEL_ARD) ? D + 1 : 2;
    c->eps = epsilon;
    c->maxB = max_chains;
    c->nslots = n_slots;
    c->maxN = max_nimp;
    c->maxNpad = (max_nimp + TB - 1) / TB * TB;
    c->mat = (size_t)c->np * c->np;
    c->slot_mode.assign(n_slots, 0);
    c->slot_valid.assign(n_slots, 0);
    const size_t B = max_chains, np = c->np;
    const size_t Bm = (parent && !full_ws) ? 1 : B;      // chains with full matrix workspaces
    const size_t own_slots = parent ? 0 : (size_t)n_slots;
    c->cached_only = parent != nullptr && !full_ws;
    int rc = APM_OK;
    auto A = [&](int r) { if (rc == APM_OK) rc = r; };
    A(dev_alloc(c, &c->dX, np * D));
    A(dev_alloc(c, &c->dy, np));
    A(dev_alloc(c, &c->dK, Bm * c->mat));
    A(dev_alloc(c, &c->dLB, Bm * c->mat));
    A(dev_alloc(c, &c->dZ, Bm * c->mat));
    if (!parent) {
        A(dev_alloc(c, &c->dSlotLK, own_slots * c->mat));
        A(dev_alloc(c, &c->dSlotLC, own_slots * c->mat));
        A(dev_alloc(c, &c->dSlotMu, own_slots * np));
        A(dev_alloc(c, &c->dSlotMt, own_slots * np));
        A(dev_alloc(c, &c->dSlotLdK, own_slots * c->nb));
        A(dev_alloc(c, &c->dSlotLdC, own_slots * c->nb));
    } else {
        c->dSlotLK = parent->dSlotLK; c->dSlotLC = parent->dSlotLC; c->dSlotMu = parent->dSlotMu;
        c->dSlotMt = parent->dSlotMt; c->dSlotLdK = parent->dSlotLdK; c->dSlotLdC = parent->dSlotLdC;
        c->root = parent;                      // slot_flags() / slot_modes() resolve to the owner
    }
    A(dev_alloc(c, &c->dLdB, B * c->nb));
    A(dev_alloc(c, &c->dInvB, Bm * (size_t)c->nb * TB * TB));
    A(dev_alloc(c, &c->dSymvDirect, B * np));
    A(dev_alloc(c, &c->dSymvPart, Bm * (size_t)c->nb * c->nb * 64));
    for (int v = 0; v < V_COUNT; v++) A(dev_alloc(c, &c->dVec[v], B * np));
    const size_t usz = B * (size_t)c->maxNpad * np;
    A(dev_alloc(c, &c->dUT, usz));
    A(dev_alloc(c, &c->dF, usz));
    A(dev_alloc(c, &c->dZf, usz));
    A(dev_alloc(c, &c->dUstage, B * (size_t)n * max_nimp));
    A(dev_alloc(c, &c->dKp, B * (size_t)(2 * D + 1)));
    A(dev_alloc(c, &c->dOut, B * 2));
    A(dev_alloc(c, &c->dLogw, B * (size_t)max_nimp));
    A(dev_alloc(c, &c->dEpDelta, B));
    A(dev_alloc(c, &c->dMaskM, B));
    A(dev_alloc(c, &c->dMaskB, B));
    A(dev_alloc(c, &c->dDoneM, B));
    A(dev_alloc(c, &c->dStatus, B));
    A(dev_alloc(c, &c->dActive, B));
    A(dev_alloc(c, &c->dIters, B));
    A(dev_alloc(c, &c->dNActive, 8));
    A(dev_alloc(c, &c->dSlotsA, B));
    A(dev_alloc(c, &c->dSlotsB, B));
    A(dev_alloc(c, &c->dFlowCounter, 8));
    A(dev_alloc(c, &c->dWork, 4));
    for (int q = 0; q < 2; q++) {
        A(dev_alloc(c, &c->dFlow2Progress[q], B * (size_t)c->nb));
        A(dev_alloc(c, &c->dFlow2Skip[q], B));
        A(dev_alloc(c, &c->dFlowYProg[q], B));
        A(dev_alloc(c, &c->dDiagPack[q], Bm * (size_t)c->nb * DP_DOUBLES));
    }
    if (rc == APM_OK) cudaMemset(c->dWork, 0, 4 * sizeof(unsigned long long));
    if (rc != APM_OK) {
        apm_destroy(c);
        return rc;
    }
    if (cudaMallocHost(&c->hKp, B * (2 * D + 1) * sizeof(double)) != cudaSuccess ||
        cudaMallocHost(&c->hOut, B * 2 * sizeof(double)) != cudaSuccess ||
        cudaMallocHost(&c->hInts, B * 4 * sizeof(int)) != cudaSuccess ||
        cudaMallocHost(&c->hNActive, 8 * sizeof(int)) != cudaSuccess) {
        set_err("apm_create: pinned host allocation failed");
        apm_destroy(c);
        return APM_ERR_NOMEM;
    }
    {
        // tensor maps of the Cholesky sources / targets
        c->tma_ok = make_matrix_tmap(&c->tmLB, c->dLB, c->np, (long long)Bm) &&
                    make_matrix_tmap(&c->tmK, c->dK, c->np, (long long)Bm) &&
                    make_matrix_tmap(&c->tmSlotLK, c->dSlotLK, c->np, (long long)n_slots) &&
                    make_matrix_tmap(&c->tmSlotLK16, c->dSlotLK, c->np, (long long)n_slots, 16) &&
                    make_matrix_tmap(&c->tmSlotLC, c->dSlotLC, c->np, (long long)n_slots);
        if (!c->tma_ok) {
            set_err("apm_create: cuTensorMapEncodeTiled failed (TMA descriptors of the Cholesky operands)");
            apm_destroy(c);
            return APM_ERR_CUDA;
        }
        int occ = 0, sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_chol_flow<true, true, 3>, CF_THREADS, CF_SMEM_BYTES) != cudaSuccess || occ < 1) occ = 1;
        c->flow2_grid = occ * sms;
        c->flow_grid_small = (occ < 2 ? occ : 2) * sms;
        if (getenv("APM_FLOW_GRID") && atoi(getenv("APM_FLOW_GRID")) > 0) c->flow2_grid = c->flow_grid_small = atoi(getenv("APM_FLOW_GRID"));
        if (getenv("APM_FLOW_SMALL_MAX")) c->flow_small_max = atoi(getenv("APM_FLOW_SMALL_MAX"));
        if (getenv("APM_NEWTON_LATE_ROUND")) c->newton_late_round = atoi(getenv("APM_NEWTON_LATE_ROUND"));
    }
    c->overlap_chol_k = getenv("APM_NO_OVERLAP") == nullptr;
    c->factored_cov = getenv("APM_EXPLICIT_COV") == nullptr;
    c->hybrid_newton = getenv("APM_NO_HYBRID_NEWTON") == nullptr;
    c->fused_vt = getenv("APM_NO_FUSED_VT") == nullptr;
    c->fused_fwd = getenv("APM_NO_FUSED_FWD") == nullptr;
    if (getenv("APM_FLOW_SPIN_NS")) c->flow_spin_ns = atoi(getenv("APM_FLOW_SPIN_NS"));
    if (getenv("APM_TRSV_CLUSTER_MAX")) c->trsv_cluster_max = atoi(getenv("APM_TRSV_CLUSTER_MAX"));
    if (getenv("APM_NEWTON_R0") && atoi(getenv("APM_NEWTON_R0")) > 0) c->newton_r0 = atoi(getenv("APM_NEWTON_R0"));
    if (getenv("APM_FNEW_THR")) c->fnew_thr = atof(getenv("APM_FNEW_THR"));
    if (getenv("APM_PRED_FACTOR") && atof(getenv("APM_PRED_FACTOR")) > 0) c->pred_factor = atof(getenv("APM_PRED_FACTOR"));
    if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_k_ready, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_lk_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_mix_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_mix_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->copy_done, cudaEventDisableTiming) != cudaSuccess) {
        set_err("apm_create: stream/event creation failed");
        apm_destroy(c);
        return APM_ERR_CUDA;
    }
    // data set: X padded with zero rows, y padded with +1
    std::vector<double> Xp(np * D, 0.0), yp(np, 1.0);
    memcpy(Xp.data(), X, sizeof(double) * (size_t)n * D);
    memcpy(yp.data(), y, sizeof(double) * n);
    for (int i = 0; i < n; i++) {
        if (!(y[i] == 1.0 || y[i] == -1.0)) {
            set_err("apm_create: targets y must be +1 / -1 (lpa.py:86 multiplies y*f)");
            apm_destroy(c);
            return APM_ERR_INVALID;
        }
    }
    cudaError_t e = cudaMemcpy(c->dX, Xp.data(), sizeof(double) * np * D, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(c->dy, yp.data(), sizeof(double) * np, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(c->dStatus, 0, sizeof(int) * B);
    if (e != cudaSuccess) {
        set_err(std::string("apm_create: upload: ") + cudaGetErrorString(e));
        apm_destroy(c);
        return APM_ERR_CUDA;
    }
    *out = c;
    return APM_OK;
}

extern "C" int apm_create(const double* X, const double* y, int n, int D, int kernel_kind, double epsilon,
                          int max_chains, int n_slots, int max_nimp, int device, apm_ctx** out) {
    return create_impl(X, y, n, D, kernel_kind, epsilon, max_chains, n_slots, max_nimp, device, nullptr, out);
}

// A second context on the parent's data set and device, with streams, workspaces and pinned staging of its own, whose
// cache slots are the parent's: apm_estimate_cached(_weights) on it may run WHILE the parent is inside
// apm_estimate_full, as long as the two calls touch different slots (the samplers' current / proposed slots of
// different chains).  Destroy it before the parent.
extern "C" int apm_create_companion(apm_ctx* parent, int max_chains, int max_nimp, apm_ctx** out) {
    if (!parent || !out || parent->root != nullptr) {
        set_err("apm_create_companion: parent must be a context created by apm_create");
        return APM_ERR_INVALID;
    }
    CU_TRY(cudaSetDevice(parent->device));
    std::vector<double> Xp((size_t)parent->np * parent->D), yp(parent->np);
    CU_TRY(cudaMemcpy(Xp.data(), parent->dX, sizeof(double) * Xp.size(), cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(yp.data(), parent->dy, sizeof(double) * yp.size(), cudaMemcpyDeviceToHost));
    return create_impl(Xp.data(), yp.data(), parent->n, parent->D, parent->kind, parent->eps, max_chains, parent->nslots,
                       max_nimp, parent->device, parent, out);
}

// internal: companion with full workspaces (see create_impl)
static int create_full_companion(apm_ctx* parent, int max_chains, int max_nimp, apm_ctx** out) {
    if (!parent || !out || parent->root != nullptr) return APM_ERR_INVALID;
    CU_TRY(cudaSetDevice(parent->device));
    std::vector<double> Xp((size_t)parent->np * parent->D), yp(parent->np);
    CU_TRY(cudaMemcpy(Xp.data(), parent->dX, sizeof(double) * Xp.size(), cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(yp.data(), parent->dy, sizeof(double) * yp.size(), cudaMemcpyDeviceToHost));
    APM_TRY(create_impl(Xp.data(), yp.data(), parent->n, parent->D, parent->kind, parent->eps, max_chains, parent->nslots,
                        max_nimp, parent->device, parent, out, true));
    apm_ctx* c = *out;     // same algorithmic settings as the parent
    c->tol = parent->tol; c->max_iters = parent->max_iters; c->approx = parent->approx;
    c->ep_tol = parent->ep_tol; c->ep_max_iters = parent->ep_max_iters; c->ep_damping = parent->ep_damping;
    return APM_OK;
}

static int not_companion(apm_ctx* c) {
    if (c && c->cached_only) {
        set_err("companion contexts (apm_create_companion) only run cached estimates");
        return APM_ERR_INVALID;
    }
    return APM_OK;
}

extern "C" int apm_destroy(apm_ctx* c) {
    if (!c) return APM_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    prof_resolve(c);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->ev_k_ready) cudaEventDestroy(c->ev_k_ready);
    if (c->ev_lk_done) cudaEventDestroy(c->ev_lk_done);
    if (c->ev_mix_fork) cudaEventDestroy(c->ev_mix_fork);
    if (c->ev_mix_join) cudaEventDestroy(c->ev_mix_join);
    if (c->copy_done) cudaEventDestroy(c->copy_done);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    for (void* p : c->allocs) cudaFree(p);
    if (c->hKp) cudaFreeHost(c->hKp);
    if (c->hOut) cudaFreeHost(c->hOut);
    if (c->hInts) cudaFreeHost(c->hInts);
    if (c->hNActive) cudaFreeHost(c->hNActive);
    delete c;
    return APM_OK;
}

extern "C" int apm_set_stream(apm_ctx* c, uint64_t s) {
    if (!c) return APM_ERR_INVALID;
    c->stream = (cudaStream_t)(uintptr_t)s;
    return APM_OK;
}
extern "C" int apm_synchronize(apm_ctx* c) {
    if (!c) return APM_ERR_INVALID;
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return APM_OK;
}
extern "C" int apm_set_overlap(apm_ctx* c, int enable) {
    if (!c) return APM_ERR_INVALID;
    c->overlap_chol_k = enable != 0;
    return APM_OK;
}
extern "C" int apm_set_newton(apm_ctx* c, double tol, int max_iters) {
    if (!c || !(tol > 0) || max_iters <= 0) return APM_ERR_INVALID;
    c->tol = tol;
    c->max_iters = max_iters;
    return APM_OK;
}
extern "C" int apm_set_approximation(apm_ctx* c, int kind, double ep_tol, int ep_max_iters, double ep_damping) {
    if (!c || (kind != 0 && kind != 1)) return APM_ERR_INVALID;
    if (kind == 1 && (!(ep_tol > 0) || ep_max_iters <= 0 || !(ep_damping > 0) || ep_damping > 1)) return APM_ERR_INVALID;
    c->approx = kind;
    if (kind == 1) {
        c->ep_tol = ep_tol;
        c->ep_max_iters = ep_max_iters;
        c->ep_damping = ep_damping;
    }
    return APM_OK;
}
extern "C" int apm_get_info(apm_ctx* c, int* n, int* D, int* n_pad, int* n_theta, int* n_slots, int* max_chains,
                            int* max_nimp) {
    if (!c) return APM_ERR_INVALID;
    if (n) *n = c->n;
    if (D) *D = c->D;
    if (n_pad) *n_pad = c->np;
    if (n_theta) *n_theta = c->P;
    if (n_slots) *n_slots = c->nslots;
    if (max_chains) *max_chains = c->maxB;
    if (max_nimp) *max_nimp = c->maxN;
    return APM_OK;
}
extern "C" int apm_profile(apm_ctx* c, int enable) {
    if (!c) return APM_ERR_INVALID;
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaStreamSynchronize(c->stream));
    prof_resolve(c);
    c->prof = enable != 0;
    return APM_OK;
}
extern "C" int apm_profile_read(apm_ctx* c, int max_entries, char* names, double* ms, int64_t* counts, int reset) {
    if (!c || !names || !ms || !counts) return -1;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    prof_resolve(c);
    int k = 0;
    for (; k < KID_COUNT && k < max_entries; k++) {
        strncpy(names + 32 * k, KID_NAMES[k], 31);
        names[32 * k + 31] = 0;
        ms[k] = c->prof_ms[k];
        counts[k] = c->prof_n[k];
        if (reset) {
            c->prof_ms[k] = 0;
            c->prof_n[k] = 0;
        }
    }
    return k;
}
extern "C" int apm_work_count(apm_ctx* c, int64_t* out, int reset) {
    if (!c || !out) return APM_ERR_INVALID;
    unsigned long long w[2] = {0, 0};
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaStreamSynchronize(c->stream));
    CU_TRY(cudaMemcpy(w, c->dWork, sizeof(w), cudaMemcpyDeviceToHost));
    out[0] = (int64_t)w[0];
    out[1] = (int64_t)w[1];
    if (reset) CU_TRY(cudaMemset(c->dWork, 0, sizeof(w)));
    return APM_OK;
}
extern "C" int64_t apm_launch_count(apm_ctx* c, int reset) {
    if (!c) return 0;
    int64_t v = c->launches;
    if (reset) c->launches = 0;
    return v;
}

// ------------------------------------------------------------------------------------------------
// building blocks
// ------------------------------------------------------------------------------------------------
static int check_B(apm_ctx* c, int B) {
    if (!c || B <= 0 || B > c->maxB) {
        set_err("batch size B out of range for this context (1..max_chains)");
        return APM_ERR_INVALID;
    }
    CU_TRY(cudaSetDevice(c->device));
    return APM_OK;
}

static int reset_status(apm_ctx* c, int B) {
    CU_TRY(cudaMemsetAsync(c->dStatus, 0, sizeof(int) * B, c->stream));
    return APM_OK;
}

// sigma = exp(theta0); ARD: tau_k = exp(theta_k); ISO: 2 tau^2 -- host libm, as the reference's libc exp
static int upload_kernel_params(apm_ctx* c, const double* theta, int B, int kind) {
    const int P = (kind == APM_KERNEL_ARD) ? c->D + 1 : 2;
    const int stride = 2 * c->D + 1;
    for (int b = 0; b < B; b++) {
        const double* th = theta + (size_t)b * P;
        double* kp = c->hKp + (size_t)b * stride;
        kp[0] = exp(th[0]);
        if (kind == APM_KERNEL_ARD) {
            for (int k = 0; k < c->D; k++) {
                kp[1 + k] = exp(th[1 + k]);
                kp[1 + c->D + k] = 1.0 / kp[1 + k];
            }
        } else {
            const double tau = exp(th[1]);
            kp[1] = 2. * (tau * tau);
            kp[2] = 1.0 / kp[1];
        }
    }
    CU_TRY(cudaMemcpyAsync(c->dKp, c->hKp, sizeof(double) * (size_t)B * stride, cudaMemcpyHostToDevice, c->stream));
    return APM_OK;
}

static int build_K(apm_ctx* c, int B, int kind, double eps) {
    KBuildParams p;
    p.X = c->dX; p.n = c->n; p.D = c->D; p.np = c->np; p.nb = c->nb;
    p.kp = c->dKp; p.kp_stride = 2 * c->D + 1;
    p.ard = (kind == APM_KERNEL_ARD); p.eps = eps;
    p.K = c->dK; p.k_bs = (long long)c->mat;
    p.ntiles = c->nb * (c->nb + 1) / 2;
    const size_t smem = (size_t)(2 * 64 * c->D + 64 * VSP + 2 * c->D + 1) * sizeof(double);
    if (smem > 96 * 1024) {
        set_err("build_K: feature dimension too large for the shared-memory staging of X");
        return APM_ERR_INVALID;
    }
    prof_begin(c, KID_BUILD_K);
    k_build_K<<<B * p.ntiles, 256, smem, c->stream>>>(p);
    return check_launch(c, "k_build_K");
}

// Batched Cholesky dst = chol(diag(scale) src diag(scale) (+ I)) for the chains with status == 0 inside `active` (null: all):
// one persistent launch of the warp-specialised TMA / mbarrier dataflow kernel (chol_flow.cuh) preceded by its queue
// initialisation.  Launches on the aux stream use the second set of queue state (they overlap the main stream's).
// work_slot >= 0: the number of chains actually factorised is added to dWork[work_slot] as well (M' factorisations).
static int run_chol(apm_ctx* c, int B, const double* src, long long src_bs, const int* src_idx, double* dst,
                    long long dst_bs, const int* dst_idx, const double* scale, int add_identity, double* logdet_parts,
                    const int* logdet_idx, int fail_code, const int* active, double* inv_out = nullptr,
                    const int* syrk_slots = nullptr, bool fwd = false, bool few_expected = false) {
    // fwd: the factorisation also solves L y = t for the right-hand side in dVec[V_T] (y -> dVec[V_S]): the forward half of
    // the Newton step's triangular solves, fused into the diagonal tasks
    // syrk_slots != null: the source is M' = P (I + L_K^T W L_K) P, accumulated on the fly from chol(K) in those slots and
    // W (dVec[V_W]) inside the factorisation's tasks (src is ignored; see chol_flow.cuh)
    cudaStream_t st = c->launch_stream ? c->launch_stream : c->stream;
    const int set = (st == c->aux_stream) ? 1 : 0;
    const CUtensorMap *tm = nullptr, *tms = nullptr;
    int m0 = 0, ms0 = 0;
    if (dst_idx) {
        tm = (dst == c->dSlotLK) ? &c->tmSlotLK : (dst == c->dSlotLC ? &c->tmSlotLC : nullptr);
    } else if (dst >= c->dLB && dst < c->dLB + (size_t)c->maxB * c->mat) {
        tm = &c->tmLB;
        m0 = (int)((dst - c->dLB) / (long long)c->mat);
    }
    // source: K (Newton / chol K), the LB buffer (M' in place) or a slot's L_C buffer (explicit covariance in place)
    if (syrk_slots) {
        tms = &c->tmK;      // (unused by the fused-source instantiation)
    } else if (src_idx) {
        tms = (src == c->dSlotLC) ? &c->tmSlotLC : (src == c->dSlotLK ? &c->tmSlotLK : nullptr);
    } else if (src >= c->dK && src < c->dK + (size_t)c->maxB * c->mat) {
        tms = &c->tmK;
        ms0 = (int)((src - c->dK) / (long long)c->mat);
    } else if (src >= c->dLB && src < c->dLB + (size_t)c->maxB * c->mat) {
        tms = &c->tmLB;
        ms0 = (int)((src - c->dLB) / (long long)c->mat);
    }
    if (!tm || !tms) {
        set_err("run_chol: source / destination buffer has no tensor map");
        return APM_ERR_INVALID;
    }
    CholFlowParams q;
    q.src = src; q.src_bs = src_bs; q.lds = c->np; q.src_idx = src_idx;
    q.dst = dst; q.dst_bs = dst_bs; q.ldd = c->np; q.dst_idx = dst_idx; q.dst_m0 = m0; q.src_m0 = ms0; q.zero = 0; q.np = c->np;
    q.scale = scale; q.scale_bs = c->np; q.add_identity = add_identity; q.nb = c->nb;
    q.logdet_parts = logdet_parts; q.logdet_stride = c->nb; q.logdet_idx = logdet_idx;
    q.inv_out = inv_out; q.inv_bs = (long long)c->nb * TB * TB;
    q.status = c->dStatus; q.fail_code = fail_code; q.active = active; q.nchains = B;
    q.counter = c->dFlowCounter + 2 * set; q.progress = c->dFlow2Progress[set]; q.list = c->dFlow2Skip[set];
    q.diagpack = c->dDiagPack[set];
    q.spin_ns = c->flow_spin_ns;
    q.lk_idx = syrk_slots; q.w = c->dVec[V_W]; q.w_bs = c->np;
    // the M' factorisations also leave V = anti-transpose of L' in the slot's L_C buffer (what the importance-sampling tail reads)
    q.vt_out = (syrk_slots && c->fused_vt) ? c->dSlotLC : nullptr; q.vt_bs = (long long)c->mat;
    q.fwd_t = fwd ? c->dVec[V_T] : nullptr; q.fwd_y = c->dVec[V_S]; q.fwd_bs = c->np; q.yprog = c->dFlowYProg[set];
    q.lt_b = c->dVec[V_B];        // <true, true>: the diagonal tasks build t' = P L_K^T b themselves
    const int total_tasks = B * c->nb * (c->nb + 1) / 2;
    prof_begin(c, KID_MISC);
    k_chol_flow_init<<<(B * c->nb + 255) / 256, 256, 0, st>>>(q.counter, q.progress, q.list, c->dStatus, active, B, c->nb, c->dWork,
                                                             syrk_slots ? c->dWork + 1 : nullptr, fwd ? q.yprog : nullptr);
    APM_TRY(check_launch(c, "k_chol_flow_init"));
    // batches whose launches are bound by the chains' critical paths run on the instantiations compiled for 2 CTAs per SM
    // (see k_chol_flow); the choice depends on the batch size only and changes no arithmetic
    // (few_expected: a late Newton round, where only stragglers are left on typical data -- a full batch would lose 1-4 %)
    const bool small = B <= c->flow_small_max || (few_expected && c->flow_small_max > 0);
    const int cap = small ? c->flow_grid_small : c->flow2_grid;
    const int grid = cap < total_tasks ? cap : total_tasks;
    prof_begin(c, KID_CHOL);
#define APM_LAUNCH_FLOW(SY, FW)                                                                                      \
    do {                                                                                                              \
        if (small) k_chol_flow<SY, FW, 2><<<grid, CF_THREADS, CF_SMEM_BYTES, st>>>(*tm, *tms, c->tmSlotLK16, q);     \
        else k_chol_flow<SY, FW, 3><<<grid, CF_THREADS, CF_SMEM_BYTES, st>>>(*tm, *tms, c->tmSlotLK16, q);           \
    } while (0)
    if (syrk_slots) {
        if (fwd) APM_LAUNCH_FLOW(true, true);
        else APM_LAUNCH_FLOW(true, false);
    } else {
        if (fwd) APM_LAUNCH_FLOW(false, true);
        else APM_LAUNCH_FLOW(false, false);
    }
#undef APM_LAUNCH_FLOW
    return check_launch(c, "k_chol_flow");
}

// out = rs * (K x) for all active chains, reading only the lower tiles of the symmetric K (lpa.py:94-95 mat-vecs)
static int run_symv(apm_ctx* c, int B, const double* x, const double* rs, double* out, const int* mask = nullptr) {
    if (!mask) mask = c->dActive;
    prof_begin(c, KID_MATVEC);
    k_symv_lower<<<dim3(c->nb, B), 256, 0, c->stream>>>(c->dK, (long long)c->mat, c->np, c->nb, x, c->np, c->dSymvDirect,
                                                        c->dSymvPart, mask, c->dStatus);
    APM_TRY(check_launch(c, "k_symv_lower"));
    prof_begin(c, KID_MATVEC);
    k_symv_reduce<<<dim3(c->nb, B), 64, 0, c->stream>>>(c->dSymvDirect, c->dSymvPart, c->nb, rs, out, c->np, mask, c->dStatus);
    return check_launch(c, "k_symv_reduce");
}

static NewtonVecs make_nv(apm_ctx* c) {
    NewtonVecs nv;
    nv.f = c->dVec[V_F]; nv.W = c->dVec[V_W]; nv.Ws = c->dVec[V_WS]; nv.bvec = c->dVec[V_B];
    nv.a = c->dVec[V_A]; nv.t = c->dVec[V_T]; nv.s = c->dVec[V_S]; nv.fnew = c->dVec[V_FNEW];
    nv.vs = c->np; nv.y = c->dy; nv.n = c->n; nv.np = c->np;
    nv.active = c->dActive; nv.iters = c->dIters; nv.status = c->dStatus; nv.n_active = c->dNActive;
    nv.tol = c->tol; nv.max_iters = c->max_iters;
    nv.done_m = nullptr; nv.mask_b = nullptr; nv.mask_m = nullptr; nv.pred_factor = c->pred_factor;
    return nv;
}

// Newton mode search (lpa.py:81-102) for chains 0..B-1 whose K sits in c->dK.  On exit f (V_F) is the
// mode, LB / Ws / a are those of the last executed iteration of each chain, dIters the iteration counts.
//
// dSlots != null (fused FULL estimate with the factored covariance): hybrid iteration.  An iteration solves
// (K^-1 + W) f_new = b.  The reference's form (B-space) is a = b - W^1/2 B^-1 W^1/2 K b, f_new = K a with
// B = I + W^1/2 K W^1/2 (one Cholesky).  The M-space form is f_new = L_K M^-1 L_K^T b with M = I + L_K^T W L_K (SYRK +
// Cholesky: twice the work) -- but chol(M) of a chain's LAST iteration is exactly what its posterior covariance needs
// (C = L_K M^-1 L_K^T with the last W, lpa.py:107-112), so a chain whose next iteration is predicted to be the last
// (k_newton_finish) runs it in M-space and skips both that iteration's chol(B) and the covariance phase: n^3/3 less.
// Both forms give the same f_new up to rounding (~1e-14 relative); a chain whose prediction fails simply iterates on
// (an unpredicted finish in B-space goes through the covariance phase as before).
// lk_pending: chol(K) is still running on the aux stream (wait for ev_lk_done before the first M-space step).
// restores the context's launch stream / Cholesky mode when a scope that redirected them ends (also on error returns)
struct StreamSwap {
    apm_ctx* c; cudaStream_t stream;
    explicit StreamSwap(apm_ctx* c_) : c(c_), stream(c_->stream) {}
    void to_aux() { c->stream = c->aux_stream; }
    void back() { c->stream = stream; }
    ~StreamSwap() { back(); }
};

// The loop is driven from the device: every kernel of a round works under the per-chain masks that k_newton_finish
// maintains (dActive; hybrid: dMaskB / dMaskM), a round whose mask is empty costs a few empty launches, and the host only
// reads the number of still-active chains after `newton_r0` rounds have been queued (then after every further round):
// one host round trip per estimate on typical data (4-5 iterations) instead of one per iteration.  In a hybrid round the
// two forms touch disjoint chains: the B-space form runs on the main stream, the M-space form beside it on the aux stream.
static int prefetch_u(apm_ctx* c, const double* u, int u_on_device, int N, int B);

static int run_newton(apm_ctx* c, int B, const int* dSlots = nullptr, bool lk_pending = false) {
    NewtonVecs nv = make_nv(c);
    const bool hybrid = dSlots != nullptr && c->factored_cov && c->hybrid_newton;
    CU_TRY(cudaMemsetAsync(nv.f, 0, sizeof(double) * (size_t)B * c->np, c->stream));
    CU_TRY(cudaMemsetAsync(c->dIters, 0, sizeof(int) * B, c->stream));
    prof_begin(c, KID_MISC);
    k_newton_init<<<(B + 255) / 256, 256, 0, c->stream>>>(c->dActive, hybrid ? c->dMaskB : nullptr, hybrid ? c->dMaskM : nullptr,
                                                          hybrid ? c->dDoneM : nullptr, c->dNActive, B);
    APM_TRY(check_launch(c, "k_newton_init"));
    if (hybrid) nv.done_m = c->dDoneM;
    const size_t trsv_c4_smem = (size_t)(c->np + 8 * 64 + 64 + 2 * 4 * 64) * sizeof(double);
    const size_t trsv_smem = (size_t)(c->np + 64 + 32 * 64) * sizeof(double);
    if (trsv_smem > 160 * 1024) {
        set_err("run_newton: n too large for the single-CTA triangular solve");
        return APM_ERR_INVALID;
    }
    // The form is chosen PER CHAIN (from its own diff only), so a chain's arithmetic never depends on its batch-mates:
    // results are bit-identical whatever the batch composition or GPU count.  A round whose active chains disagree runs
    // both forms, each under its mask (dMaskB / dMaskM, written by k_newton_finish).
    c->newton_b_finishers = false;
    c->newton_b_finisher_count = 0;
    const int* maskB = c->dActive;
    const int* maskM = nullptr;
    if (hybrid) {
        nv.mask_b = c->dMaskB; nv.mask_m = c->dMaskM;
        maskB = c->dMaskB; maskM = c->dMaskM;
    }
    // mat-vec-free B-space step (k_newton_prep / k_trsv2 / k_fnew_from_s): no K-sized read besides the Cholesky itself
    const bool matfree = c->fnew_thr > 0;
    NewtonVecs nvB = nv, nvM = nv;     // k_trsv2 skips the chains outside nv.active
    nvB.active = const_cast<int*>(maskB);
    nvM.active = const_cast<int*>(maskM);
    const bool two_streams = c->overlap_chol_k && c->aux_stream != nullptr;
    int n_act = B;
    for (int it = 0; it < c->max_iters; it++) {
        prof_begin(c, KID_NEWTON_VEC);
        k_newton_prep<<<B, 256, 0, c->stream>>>(nv, matfree ? 1 : 0);
        APM_TRY(check_launch(c, "k_newton_prep"));
        // no chain can be in M-space in the first round (the prediction needs a diff)
        const bool m_form = hybrid && it > 0;
        StreamSwap swap(c);
        cudaStream_t main_stream = c->stream;
        const bool fork = m_form && two_streams;
        if (fork) {
            CU_TRY(cudaEventRecord(c->ev_mix_fork, main_stream));
            CU_TRY(cudaStreamWaitEvent(c->aux_stream, c->ev_mix_fork, 0));
        }
        {
            // t = Ws * (K b)                                           (lpa.py:94  W_sqrt_K.dot(b)); mat-vec-free: t = b / Ws
            if (!matfree) APM_TRY(run_symv(c, B, nv.bvec, nv.Ws, nv.t, maskB));
            // L = chol(I + Ws K Ws)                                    (lpa.py:91-92)
            APM_TRY(run_chol(c, B, c->dK, (long long)c->mat, nullptr, c->dLB, (long long)c->mat, nullptr, nv.Ws, 1, c->dLdB,
                             nullptr, APM_CHAIN_CHOL_B, maskB, c->dInvB, nullptr, c->fused_fwd, it >= c->newton_late_round));
            // s = L^-T L^-1 t ; a = b - Ws s                           (lpa.py:94)
            prof_begin(c, KID_TRSV);
            if (c->fused_fwd && B <= c->trsv_cluster_max)     // a batch of about one chain per SM or less: latency-bound, 4 CTAs per chain
                k_trsv_back_c4<<<4 * B, 256, trsv_c4_smem, c->stream>>>(c->dLB, (long long)c->mat, c->np, c->nb, c->dInvB,
                                                                        (long long)c->nb * TB * TB, nvB, B);
            else if (c->fused_fwd)
                k_trsv2<true><<<B, 256, trsv_smem, c->stream>>>(c->dLB, (long long)c->mat, c->np, c->nb, c->dInvB,
                                                                (long long)c->nb * TB * TB, nvB);
            else
                k_trsv2<false><<<B, 256, trsv_smem, c->stream>>>(c->dLB, (long long)c->mat, c->np, c->nb, c->dInvB,
                                                                 (long long)c->nb * TB * TB, nvB);
            APM_TRY(check_launch(c, "k_trsv2"));
            // f_new = K a                                              (lpa.py:95)
            if (matfree) {
                prof_begin(c, KID_MATVEC);
                k_fnew_from_s<<<B, 256, 0, c->stream>>>(c->dK, (long long)c->mat, c->np, nvB, c->fnew_thr, 1);
                APM_TRY(check_launch(c, "k_fnew_from_s"));
            } else {
                APM_TRY(run_symv(c, B, nv.a, nullptr, nv.fnew, maskB));
            }
        }
        if (m_form) {
            if (fork) swap.to_aux();
            if (lk_pending) {
                CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_lk_done, 0));
                lk_pending = false;
            }
            // t' = reversed L_K^T b: with the fused forward substitution the factorisation's diagonal tasks accumulate it from the
            // L_K boxes they stream for M'; otherwise by its own kernel
            if (!c->fused_fwd) {
                prof_begin(c, KID_MATVEC);
                k_lt_matvec<<<dim3(c->nb, B), 256, 0, c->stream>>>(c->dSlotLK, (long long)c->mat, dSlots, c->np, c->nb, nv.bvec, c->np,
                                                                    nv.t, c->np, nullptr, c->dStatus, maskM, 1);
                APM_TRY(check_launch(c, "k_lt_matvec"));
            }
            // L' = chol(M'), M' = P (I + L_K^T W L_K) P built inside the factorisation from L_K and W (never stored)
            APM_TRY(run_chol(c, B, nullptr, 0, nullptr, c->dLB, (long long)c->mat, nullptr, nullptr, 0, c->dLdB,
                             nullptr, APM_CHAIN_CHOL_C, maskM, c->dInvB, dSlots, c->fused_fwd, it >= c->newton_late_round));
            // s' = M'^-1 t'
            prof_begin(c, KID_TRSV);
            if (c->fused_fwd && B <= c->trsv_cluster_max)     // a batch of about one chain per SM or less: latency-bound, 4 CTAs per chain
                k_trsv_back_c4<<<4 * B, 256, trsv_c4_smem, c->stream>>>(c->dLB, (long long)c->mat, c->np, c->nb, c->dInvB,
                                                                        (long long)c->nb * TB * TB, nvM, B);
            else if (c->fused_fwd)
                k_trsv2<true><<<B, 256, trsv_smem, c->stream>>>(c->dLB, (long long)c->mat, c->np, c->nb, c->dInvB,
                                                                (long long)c->nb * TB * TB, nvM);
            else
                k_trsv2<false><<<B, 256, trsv_smem, c->stream>>>(c->dLB, (long long)c->mat, c->np, c->nb, c->dInvB,
                                                                 (long long)c->nb * TB * TB, nvM);
            APM_TRY(check_launch(c, "k_trsv2"));
            // mu~ = reversed s' (-> slot), f_new = L_K mu~
            prof_begin(c, KID_MATVEC);
            k_l_matvec_rev<<<dim3(c->np / 32, B), 256, 0, c->stream>>>(c->dSlotLK, (long long)c->mat, dSlots, c->np, c->np, nv.s,
                                                                       c->np, nv.fnew, c->np, c->dSlotMt, c->np, dSlots, c->dStatus,
                                                                       maskM);
            APM_TRY(check_launch(c, "k_l_matvec_rev"));
            swap.back();
            if (fork) {
                CU_TRY(cudaEventRecord(c->ev_mix_join, c->aux_stream));
                CU_TRY(cudaStreamWaitEvent(main_stream, c->ev_mix_join, 0));
            }
        }
        prof_begin(c, KID_NEWTON_VEC);
        k_newton_finish<<<B, 256, 0, c->stream>>>(nv);
        APM_TRY(check_launch(c, "k_newton_finish"));
        if (it + 1 < c->newton_r0 && it + 1 < c->max_iters) continue;     // keep queueing: no host round trip yet
        CU_TRY(cudaMemcpyAsync(c->hNActive, c->dNActive, 3 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        if (c->pend_u) {       // the GPU is busy with the queued rounds: now is the time to bring u over
            APM_TRY(prefetch_u(c, c->pend_u, 0, c->pend_N, c->pend_B));
            c->pend_u = nullptr;
        }
        CU_TRY(cudaStreamSynchronize(c->stream));
        n_act = c->hNActive[0];
        c->newton_b_finisher_count = hybrid ? c->hNActive[2] : 0;   // (cumulative) finished in a B-space round: need the covariance phase
        if (n_act <= 0) break;
    }
    c->newton_b_finishers = !hybrid || c->newton_b_finisher_count > 0;
    return APM_OK;
}

// Z L_B^T = K W^1/2 (lpa.py:111 as a right triangular solve): Z = K W^1/2 L_B^{-T} into dZ, for chains in `active`
static int run_trsm_z(apm_ctx* c, int B, const int* active) {
    TrsmParams t;
    t.R = c->dK; t.r_bs = (long long)c->mat; t.ldr = c->np; t.r_idx = nullptr;
    t.cs = c->dVec[V_WS]; t.cs_bs = c->np;
    t.X = c->dZ; t.x_bs = (long long)c->mat; t.ldx = c->np;
    t.L = c->dLB; t.l_bs = (long long)c->mat; t.ldl = c->np; t.l_idx = nullptr;
    t.nb = c->nb; t.row_blocks = c->nb;
    t.status = c->dStatus; t.active = active;
    prof_begin(c, KID_TRSM);
    k_trsm_rows<<<B * c->nb, TILE_THREADS, TILE_SMEM_BYTES, c->stream>>>(t);
    return check_launch(c, "k_trsm_rows");
}

// Parallel EP (extension; the ep_approximation restatement under oracle/) for chains 0..B-1 whose K sits in c->dK.  On exit
// f (V_F) = posterior mean, Ws = sqrt(tau~), bvec = nu~, LB / Z those of each chain's last iteration, dIters the
// iteration counts.  Per iteration: site update (O(n)), then Sigma and mu from scratch exactly as the Newton step does
// its solve (B = I + S^1/2 K S^1/2, a = nu~ - S^1/2 B^-1 S^1/2 K nu~, mu = K a) plus diag(Sigma) from Z.
static int run_ep(apm_ctx* c, int B) {
    NewtonVecs nv = make_nv(c);
    nv.tol = c->ep_tol;
    nv.max_iters = c->ep_max_iters;
    EpVecs ev;
    ev.s2 = c->dVec[V_S2]; ev.delta = c->dEpDelta; ev.damping = c->ep_damping;
    CU_TRY(cudaMemsetAsync(nv.f, 0, sizeof(double) * (size_t)B * c->np, c->stream));
    CU_TRY(cudaMemsetAsync(nv.W, 0, sizeof(double) * (size_t)B * c->np, c->stream));
    CU_TRY(cudaMemsetAsync(nv.bvec, 0, sizeof(double) * (size_t)B * c->np, c->stream));
    CU_TRY(cudaMemsetAsync(c->dIters, 0, sizeof(int) * B, c->stream));
    prof_begin(c, KID_MISC);
    k_fill_int<<<(B + 255) / 256, 256, 0, c->stream>>>(c->dActive, 1, B);
    APM_TRY(check_launch(c, "k_fill_int"));
    prof_begin(c, KID_MISC);
    k_fill_int<<<1, 32, 0, c->stream>>>(c->dNActive, B, 1);
    APM_TRY(check_launch(c, "k_fill_int"));
    prof_begin(c, KID_NEWTON_VEC);
    k_ep_init<<<B, 256, 0, c->stream>>>(c->dK, (long long)c->mat, c->np, nv, ev);
    APM_TRY(check_launch(c, "k_ep_init"));
    const size_t trsv_smem = (size_t)(c->np + 64 + 8 * 64) * sizeof(double);
    if (trsv_smem > 160 * 1024) {
        set_err("run_ep: n too large for the single-CTA triangular solve");
        return APM_ERR_INVALID;
    }
    for (int it = 0; it < c->ep_max_iters; it++) {
        prof_begin(c, KID_NEWTON_VEC);
        k_ep_sites<<<B, 256, 0, c->stream>>>(nv, ev);
        APM_TRY(check_launch(c, "k_ep_sites"));
        APM_TRY(run_symv(c, B, nv.bvec, nv.Ws, nv.t));                           // t = S^1/2 K nu~
        APM_TRY(run_chol(c, B, c->dK, (long long)c->mat, nullptr, c->dLB, (long long)c->mat, nullptr, nv.Ws, 1, c->dLdB,
                         nullptr, APM_CHAIN_CHOL_B, c->dActive, c->dInvB));
        prof_begin(c, KID_TRSV);
        k_trsv2<false><<<B, 256, trsv_smem, c->stream>>>(c->dLB, (long long)c->mat, c->np, c->nb, c->dInvB,
                                                         (long long)c->nb * TB * TB, nv);   // a = nu~ - S^1/2 B^-1 t
        APM_TRY(check_launch(c, "k_trsv2"));
        APM_TRY(run_symv(c, B, nv.a, nullptr, nv.fnew));                         // mu = K a
        APM_TRY(run_trsm_z(c, B, c->dActive));
        prof_begin(c, KID_NEWTON_VEC);
        k_ep_diag_sigma<<<dim3(c->np / 32, B), 256, 0, c->stream>>>(c->dK, (long long)c->mat, c->dZ, (long long)c->mat, c->np, nv, ev);
        APM_TRY(check_launch(c, "k_ep_diag_sigma"));
        prof_begin(c, KID_NEWTON_VEC);
        k_ep_finish<<<B, 256, 0, c->stream>>>(nv, ev);
        APM_TRY(check_launch(c, "k_ep_finish"));
        CU_TRY(cudaMemcpyAsync(c->hNActive, c->dNActive, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
        if (c->hNActive[0] <= 0) break;
    }
    return APM_OK;
}

// C = K - Z Z^T with Z L_B^T = K W^1/2  (lpa.py:111-112), lower tiles into dst
static int run_covariance(apm_ctx* c, int B, double* dst, long long dst_bs, const int* dst_idx) {
    TrsmParams t;
    t.R = c->dK; t.r_bs = (long long)c->mat; t.ldr = c->np; t.r_idx = nullptr;
    t.cs = c->dVec[V_WS]; t.cs_bs = c->np;
    t.X = c->dZ; t.x_bs = (long long)c->mat; t.ldx = c->np;
    t.L = c->dLB; t.l_bs = (long long)c->mat; t.ldl = c->np; t.l_idx = nullptr;
    t.nb = c->nb; t.row_blocks = c->nb;
    t.status = c->dStatus; t.active = nullptr;
    prof_begin(c, KID_TRSM);
    k_trsm_rows<<<B * c->nb, TILE_THREADS, TILE_SMEM_BYTES, c->stream>>>(t);
    APM_TRY(check_launch(c, "k_trsm_rows"));
    SyrkParams s;
    s.S = c->dK; s.s_bs = (long long)c->mat; s.lds = c->np;
    s.Z = c->dZ; s.z_bs = (long long)c->mat; s.ldz = c->np;
    s.C = dst; s.c_bs = dst_bs; s.ldc = c->np; s.c_idx = dst_idx;
    s.nb = c->nb; s.ntiles = c->nb * (c->nb + 1) / 2;
    s.status = c->dStatus;
    prof_begin(c, KID_SYRK);
    k_syrk_sub<<<B * s.ntiles, TILE_THREADS, TILE_SMEM_BYTES, c->stream>>>(s);
    return check_launch(c, "k_syrk_sub");
}

// chol(C) without forming C (see tile_engine.cuh "Factored posterior covariance"): needs chol(K) in the slots, W^1/2 and
// `a` of the last Newton / EP step; writes V (slot mode 1), mu~ and the partial log-dets of C into the slots.
static int run_covariance_factored(apm_ctx* c, int B, const int* dSlots, const int* done_m = nullptr) {
    // chains whose last Newton iteration ran in M-space (done_m) already have L' = chol(M') in dLB and mu~ in the slot
    const int* todo = nullptr;
    const bool need_cov = !done_m || c->newton_b_finishers;
    if (done_m && need_cov) {
        prof_begin(c, KID_MISC);
        k_mask_not<<<(B + 255) / 256, 256, 0, c->stream>>>(done_m, c->dMaskB, B);
        APM_TRY(check_launch(c, "k_mask_not"));
        todo = c->dMaskB;
    }
    if (need_cov) {
        // L' = chol(M'), M' = P (I + L_K^T W L_K) P from L_K and W (M' has eigenvalues >= 1: cannot fail for finite input)
        APM_TRY(run_chol(c, B, nullptr, 0, nullptr, c->dLB, (long long)c->mat, nullptr, nullptr, 0, c->dLdB,
                         nullptr, APM_CHAIN_CHOL_C, todo, nullptr, dSlots));
    }
    // V = anti-transpose of L' is already in the slot's L_C buffer: every M' factorisation (k_chol_flow<true>) stores its tiles
    // a second time in that form.  mu~ = L_K^T a.  The importance-sampling tail works with (L_K, V, mu, mu~) directly, so the
    // n^3/3 triangular solve for the explicit L_C = L_K V^-1 is only run if somebody asks for C_chol (slot_make_explicit).
    if (!c->fused_vt) {
        dim3 ag(c->np / 32, c->np / 32, B), ab(32, 8);
        prof_begin(c, KID_TRANSPOSE);
        k_antitranspose<<<ag, ab, 0, c->stream>>>(c->dLB, (long long)c->mat, nullptr, c->dSlotLC, (long long)c->mat, dSlots, c->np,
                                                 c->dStatus);
        APM_TRY(check_launch(c, "k_antitranspose"));
    }
    if (need_cov) {
        prof_begin(c, KID_MATVEC);
        k_lt_matvec<<<dim3(c->nb, B), 256, 0, c->stream>>>(c->dSlotLK, (long long)c->mat, dSlots, c->np, c->nb, c->dVec[V_A], c->np,
                                                            c->dSlotMt, c->np, dSlots, c->dStatus, todo, 0);
        APM_TRY(check_launch(c, "k_lt_matvec"));
    }
    prof_begin(c, KID_MISC);
    k_logdet_combine<<<B, 256, 0, c->stream>>>(c->dSlotLdK, c->dLdB, c->dSlotLdC, dSlots, c->nb, c->dStatus);
    return check_launch(c, "k_logdet_combine");
}

// slot mode 1 -> 0: L_C = L_K V^-1 by the reversed right triangular solve (k_trsm_rev) with L' = anti-transpose of V.
// Uses chain 0 of the scratch matrices (the API is host-synchronous: nothing else is in flight).
static int slot_make_explicit(apm_ctx* c, int slot) {
    if (slot_modes(c)[slot] != 1) return APM_OK;
    c->hInts[0] = slot;
    CU_TRY(cudaMemcpyAsync(c->dSlotsB, c->hInts, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    dim3 ag(c->np / 32, c->np / 32, 1), ab(32, 8);
    prof_begin(c, KID_TRANSPOSE);
    k_antitranspose<<<ag, ab, 0, c->stream>>>(c->dSlotLC, (long long)c->mat, c->dSlotsB, c->dLB, (long long)c->mat, nullptr, c->np,
                                             nullptr);
    APM_TRY(check_launch(c, "k_antitranspose"));
    CU_TRY(cudaMemsetAsync(c->dFlow2Skip[0], 0, sizeof(int), c->stream));   // a zero status word for the single pseudo-chain
    TrsmRevParams t;
    t.LK = c->dSlotLK; t.lk_bs = (long long)c->mat; t.ldk = c->np; t.lk_idx = c->dSlotsB;
    t.Lp = c->dLB; t.lp_bs = (long long)c->mat; t.ldp = c->np;
    t.X = c->dZ; t.x_bs = (long long)c->mat; t.ldx = c->np;
    t.LC = c->dSlotLC; t.lc_bs = (long long)c->mat; t.ldc = c->np; t.lc_idx = c->dSlotsB;
    t.nb = c->nb;
    t.status = c->dFlow2Skip[0];
    prof_begin(c, KID_TRSM);
    k_trsm_rev<<<c->nb, TILE_THREADS, TILE_SMEM_BYTES, c->stream>>>(t);
    APM_TRY(check_launch(c, "k_trsm_rev"));
    CU_TRY(cudaStreamSynchronize(c->stream));
    slot_modes(c)[slot] = 0;
    return APM_OK;
}

// bring u (reference layout [B][n][N]) into UT [B][Npad][np]
static int stage_u(apm_ctx* c, const double* u, int u_on_device, int N, int B) {
    if (N <= 0 || N > c->maxN) {
        set_err("N (importance samples) out of range for this context");
        return APM_ERR_INVALID;
    }
    const double* du = u;
    if (!u_on_device) {
        if (c->u_staged) {   // upload was started on the copy stream by prefetch_u
            CU_TRY(cudaStreamWaitEvent(c->stream, c->copy_done, 0));
            c->u_staged = false;
        } else {
            CU_TRY(cudaMemcpyAsync(c->dUstage, u, sizeof(double) * (size_t)B * c->n * N, cudaMemcpyHostToDevice, c->stream));
        }
        du = c->dUstage;
    }
    const int Npad = (N + TB - 1) / TB * TB;
    dim3 grid(c->np / 32, Npad / 32, B), block(32, 8);
    prof_begin(c, KID_TRANSPOSE);
    k_transpose_u<<<grid, block, 0, c->stream>>>(du, (long long)c->n * N, c->n, N, c->dUT, (long long)Npad * c->np, c->np,
                                                 Npad);
    return check_launch(c, "k_transpose_u");
}

// drop a pending prefetch (a previous call returned early): the staging buffer must be quiescent before re-use
static void cancel_prefetch(apm_ctx* c) {
    if (c->u_staged) {
        cudaStreamSynchronize(c->copy_stream);
        c->u_staged = false;
    }
}

// start the host->device copy of u on the copy stream so that it overlaps the factorisations; stage_u waits
static int prefetch_u(apm_ctx* c, const double* u, int u_on_device, int N, int B) {
    if (u_on_device || N <= 0 || N > c->maxN) return APM_OK;
    CU_TRY(cudaMemcpyAsync(c->dUstage, u, sizeof(double) * (size_t)B * c->n * N, cudaMemcpyHostToDevice, c->copy_stream));
    CU_TRY(cudaEventRecord(c->copy_done, c->copy_stream));
    c->u_staged = true;
    return APM_OK;
}

// the O(n^2 N) tail (estimators.py:221-241) for chains whose caches sit in slots dSlots[b]
// mode 0: importance sampling from the posterior approximation; 1: prior Monte Carlo.  factored (mode 0 only): the slots
// hold V and mu~ instead of chol(C) (slot mode 1).
static int run_is_tail(apm_ctx* c, int N, int B, const int* dSlots, double* d_logml, double* d_logw, int mode, bool factored = false) {
    const int Npad = (N + TB - 1) / TB * TB;
    const int rblocks = Npad / TB;
    const long long ubs = (long long)Npad * c->np;
    if (mode == 0 && factored) {
        // W V^T = U^T  (w_s = V^-T u_s), rows = samples                       [same flops as estimators.py:225's solve]
        TrsmParams t;
        t.R = c->dUT; t.r_bs = ubs; t.ldr = c->np; t.r_idx = nullptr;
        t.cs = nullptr; t.cs_bs = 0;
        t.X = c->dZf; t.x_bs = ubs; t.ldx = c->np;
        t.L = c->dSlotLC; t.l_bs = (long long)c->mat; t.ldl = c->np; t.l_idx = dSlots;
        t.nb = c->nb; t.row_blocks = rblocks;
        t.status = c->dStatus; t.active = nullptr;
        prof_begin(c, KID_TRSM);
        k_trsm_rows<<<B * rblocks, TILE_THREADS, TILE_SMEM_BYTES, c->stream>>>(t);
        APM_TRY(check_launch(c, "k_trsm_rows"));
    }
    GemmTriParams g;
    // F = mu + U^T L_C^T (estimators.py:223); factored: F = mu + W L_K^T; prior MC: F = U^T L_K^T (estimators.py:323)
    g.UT = (mode == 0 && factored) ? c->dZf : c->dUT; g.u_bs = ubs; g.ldu = c->np;
    g.L = (mode == 0 && !factored) ? c->dSlotLC : c->dSlotLK; g.l_bs = (long long)c->mat; g.ldl = c->np; g.l_idx = dSlots;
    g.mu = (mode == 0) ? c->dSlotMu : nullptr; g.mu_bs = c->np; g.mu_idx = dSlots;
    g.F = c->dF; g.f_bs = ubs; g.ldf = c->np;
    g.nb = c->nb; g.row_blocks = rblocks;
    g.status = c->dStatus;
    prof_begin(c, KID_GEMM_TRI);
    k_gemm_tri<<<B * c->nb * rblocks, TILE_THREADS, TILE_SMEM_BYTES, c->stream>>>(g);
    APM_TRY(check_launch(c, "k_gemm_tri"));
    if (mode == 0 && !factored) {
        TrsmParams t;                                                          // Zf = L_K^-1 f_s (estimators.py:225)
        t.R = c->dF; t.r_bs = ubs; t.ldr = c->np; t.r_idx = nullptr;
        t.cs = nullptr; t.cs_bs = 0;
        t.X = c->dZf; t.x_bs = ubs; t.ldx = c->np;
        t.L = c->dSlotLK; t.l_bs = (long long)c->mat; t.ldl = c->np; t.l_idx = dSlots;
        t.nb = c->nb; t.row_blocks = rblocks;
        t.status = c->dStatus; t.active = nullptr;
        prof_begin(c, KID_TRSM);
        k_trsm_rows<<<B * rblocks, TILE_THREADS, TILE_SMEM_BYTES, c->stream>>>(t);
        APM_TRY(check_launch(c, "k_trsm_rows"));
    }
    EpilogueParams e;
    e.F = c->dF; e.Zf = c->dZf; e.UT = c->dUT; e.bs = ubs; e.ld = c->np;
    e.y = c->dy; e.n = c->n; e.N = N;
    e.logdetK = c->dSlotLdK; e.logdetC = c->dSlotLdC; e.ld_stride = c->nb; e.nb = c->nb; e.slot_idx = dSlots;
    e.status = c->dStatus;
    e.logml = d_logml; e.logw = d_logw ? d_logw : c->dLogw;
    e.mode = mode;
    e.mt = (mode == 0 && factored) ? c->dSlotMt : nullptr; e.mt_bs = c->np;   // L_K^-1 f_s = mu~ + w_s
    prof_begin(c, KID_EPILOGUE);
    k_is_logw<<<dim3((N + 7) / 8, B), 256, 0, c->stream>>>(e);
    APM_TRY(check_launch(c, "k_is_logw"));
    prof_begin(c, KID_EPILOGUE);
    k_is_epilogue<<<B, 256, 0, c->stream>>>(e);
    return check_launch(c, "k_is_epilogue");
}

static int upload_slots(apm_ctx* c, const int* slots, int B, int* dSlots, bool require_valid) {
    for (int b = 0; b < B; b++) {
        if (slots[b] < 0 || slots[b] >= c->nslots) {
            set_err("slot index out of range");
            return APM_ERR_INVALID;
        }
        if (require_valid && !slot_flags(c)[slots[b]]) {
            set_err("slot " + std::to_string(slots[b]) + " holds no valid cache");
            return APM_ERR_INVALID;
        }
        c->hInts[b] = slots[b];
    }
    CU_TRY(cudaMemcpyAsync(dSlots, c->hInts, sizeof(int) * B, cudaMemcpyHostToDevice, c->stream));
    // hInts is re-used by the caller only after a stream sync
    CU_TRY(cudaStreamSynchronize(c->stream));
    return APM_OK;
}

static int fetch_results(apm_ctx* c, int B, double* out_d, int n_out, double* out_h, int* iters_h, int iters_add,
                         int* status_h) {
    if (out_h) CU_TRY(cudaMemcpyAsync(c->hOut, out_d, sizeof(double) * (size_t)B * n_out, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaMemcpyAsync(c->hInts, c->dIters, sizeof(int) * B, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaMemcpyAsync(c->hInts + B, c->dStatus, sizeof(int) * B, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    if (c->prof) prof_resolve(c);
    if (out_h) memcpy(out_h, c->hOut, sizeof(double) * (size_t)B * n_out);
    for (int b = 0; b < B; b++) {
        if (iters_h) iters_h[b] = c->hInts[b] + iters_add;
        if (status_h) status_h[b] = c->hInts[B + b];
    }
    return APM_OK;
}

// ------------------------------------------------------------------------------------------------
// C ABI: kernels
// ------------------------------------------------------------------------------------------------
extern "C" int apm_kernel_build(apm_ctx* c, const double* theta, int B, int kernel_kind, double epsilon, double* K_out,
                                int K_on_device) {
    APM_TRY(not_companion(c));
    APM_TRY(check_B(c, B));
    if (!theta || !K_out) return APM_ERR_INVALID;
    const int kind = kernel_kind < 0 ? c->kind : kernel_kind;
    const double eps = epsilon < 0 ? c->eps : epsilon;
    APM_TRY(upload_kernel_params(c, theta, B, kind));
    APM_TRY(build_K(c, B, kind, eps));
    const size_t n = c->n;
    for (int b = 0; b < B; b++) {
        CU_TRY(cudaMemcpy2DAsync(K_out + (size_t)b * n * n, n * sizeof(double), c->dK + (size_t)b * c->mat,
                                 (size_t)c->np * sizeof(double), n * sizeof(double), n,
                                 K_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
    }
    CU_TRY(cudaStreamSynchronize(c->stream));
    return APM_OK;
}

extern "C" int apm_kernel_grad(apm_ctx* c, const double* theta, int B, int kernel_kind, double* dK_out, int dK_on_device) {
    APM_TRY(not_companion(c));
    APM_TRY(check_B(c, B));
    if (!theta || !dK_out) return APM_ERR_INVALID;
    const int kind = kernel_kind < 0 ? c->kind : kernel_kind;
    const int P = (kind == APM_KERNEL_ARD) ? c->D + 1 : 2;
    APM_TRY(upload_kernel_params(c, theta, B, kind));
    const size_t count = (size_t)B * P * c->n * c->n;
    double* dst = dK_out;
    if (!dK_on_device) {
        cudaError_t e = cudaMalloc(&dst, count * sizeof(double));
        if (e != cudaSuccess) {
            set_err(std::string("apm_kernel_grad: staging buffer: ") + cudaGetErrorString(e));
            return APM_ERR_NOMEM;
        }
    }
    KGradParams p;
    p.X = c->dX; p.n = c->n; p.D = c->D; p.nb = c->nb;
    p.kp = c->dKp; p.kp_stride = 2 * c->D + 1;
    p.ard = (kind == APM_KERNEL_ARD); p.P = P;
    p.dK = dst;
    const size_t smem = (size_t)(2 * 64 * c->D + 2 * c->D + 1) * sizeof(double);
    int rc = APM_OK;
    if (smem > 96 * 1024) {
        set_err("apm_kernel_grad: feature dimension too large for the shared-memory staging of X");
        rc = APM_ERR_INVALID;
    } else {
        prof_begin(c, KID_BUILD_K);
        k_build_dK<<<B * c->nb * c->nb, 256, smem, c->stream>>>(p);
        rc = check_launch(c, "k_build_dK");
    }
    cudaError_t e = cudaSuccess;
    if (rc == APM_OK && !dK_on_device) e = cudaMemcpyAsync(dK_out, dst, count * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (!dK_on_device) cudaFree(dst);
    if (rc == APM_OK && e != cudaSuccess) {
        set_err(std::string("apm_kernel_grad: ") + cudaGetErrorString(e));
        rc = APM_ERR_CUDA;
    }
    return rc;
}

static int import_matrices(apm_ctx* c, const double* M, int on_device, int B, double* dst, long long dst_bs) {
    const size_t n = c->n;
    for (int b = 0; b < B; b++) {
        CU_TRY(cudaMemcpy2DAsync(dst + (size_t)b * dst_bs, (size_t)c->np * sizeof(double), M + (size_t)b * n * n,
                                 n * sizeof(double), n * sizeof(double), n,
                                 on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    }
    if (c->np > c->n) {
        dim3 grid(64, B);
        prof_begin(c, KID_MISC);
        k_pad_identity<<<grid, 256, 0, c->stream>>>(dst, dst_bs, c->n, c->np);
        APM_TRY(check_launch(c, "k_pad_identity"));
    }
    return APM_OK;
}

extern "C" int apm_laplace(apm_ctx* c, const double* K, int K_on_device, int B, int calc_cov, int calc_lml, double* f_out,
                           double* C_out, int C_on_device, double* lml_out, int* cubic_ops_out, int* chain_status) {
    APM_TRY(not_companion(c));
    APM_TRY(check_B(c, B));
    if (!K) return APM_ERR_INVALID;
    APM_TRY(reset_status(c, B));
    APM_TRY(import_matrices(c, K, K_on_device, B, c->dK, (long long)c->mat));
    APM_TRY(run_newton(c, B));
    NewtonVecs nv = make_nv(c);
    if (calc_lml) {
        prof_begin(c, KID_NEWTON_VEC);
        k_laplace_lml<<<B, 256, 0, c->stream>>>(nv, c->dLdB, c->nb, c->nb, c->dOut);
        APM_TRY(check_launch(c, "k_laplace_lml"));
    }
    if (calc_cov && C_out) {
        // C (lower tiles) -> LB buffer (chol(B) is dead after the triangular solve), then a dense symmetric export
        APM_TRY(run_covariance(c, B, c->dLB, (long long)c->mat, nullptr));
        const size_t n = c->n;
        for (int b = 0; b < B; b++) {
            dim3 grid((c->n + 255) / 256, c->n);
            prof_begin(c, KID_MISC);
            k_export_lower<<<grid, 256, 0, c->stream>>>(c->dLB + (size_t)b * c->mat, c->np, c->n, c->dZ + (size_t)b * c->mat, 1);
            APM_TRY(check_launch(c, "k_export_lower"));
            CU_TRY(cudaMemcpyAsync(C_out + (size_t)b * n * n, c->dZ + (size_t)b * c->mat, sizeof(double) * n * n,
                                   C_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
        }
    }
    if (f_out) {
        CU_TRY(cudaMemcpy2DAsync(f_out, (size_t)c->n * sizeof(double), nv.f, (size_t)c->np * sizeof(double),
                                 (size_t)c->n * sizeof(double), B, cudaMemcpyDeviceToHost, c->stream));
    }
    return fetch_results(c, B, c->dOut, 1, calc_lml ? lml_out : nullptr, cubic_ops_out, calc_cov ? 1 : 0, chain_status);
}

extern "C" int apm_ep(apm_ctx* c, const double* K, int K_on_device, int B, int calc_cov, double* f_out, double* C_out,
                      int C_on_device, double* nu_out, double* tau_out, int* cubic_ops_out, int* chain_status) {
    APM_TRY(not_companion(c));
    APM_TRY(check_B(c, B));
    if (!K) return APM_ERR_INVALID;
    APM_TRY(reset_status(c, B));
    APM_TRY(import_matrices(c, K, K_on_device, B, c->dK, (long long)c->mat));
    APM_TRY(run_ep(c, B));
    NewtonVecs nv = make_nv(c);
    if (calc_cov && C_out) {
        // Sigma = K - Z Z^T with the Z of each chain's last iteration (lower tiles -> LB buffer, dense symmetric export)
        SyrkParams s;
        s.S = c->dK; s.s_bs = (long long)c->mat; s.lds = c->np;
        s.Z = c->dZ; s.z_bs = (long long)c->mat; s.ldz = c->np;
        s.C = c->dLB; s.c_bs = (long long)c->mat; s.ldc = c->np; s.c_idx = nullptr;
        s.nb = c->nb; s.ntiles = c->nb * (c->nb + 1) / 2;
        s.status = c->dStatus;
        prof_begin(c, KID_SYRK);
        k_syrk_sub<<<B * s.ntiles, TILE_THREADS, TILE_SMEM_BYTES, c->stream>>>(s);
        APM_TRY(check_launch(c, "k_syrk_sub"));
        const size_t n = c->n;
        for (int b = 0; b < B; b++) {
            dim3 grid((c->n + 255) / 256, c->n);
            prof_begin(c, KID_MISC);
            k_export_lower<<<grid, 256, 0, c->stream>>>(c->dLB + (size_t)b * c->mat, c->np, c->n, c->dZ + (size_t)b * c->mat, 1);
            APM_TRY(check_launch(c, "k_export_lower"));
            CU_TRY(cudaMemcpyAsync(C_out + (size_t)b * n * n, c->dZ + (size_t)b * c->mat, sizeof(double) * n * n,
                                   C_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
        }
    }
    const size_t row = (size_t)c->n * sizeof(double), pitch = (size_t)c->np * sizeof(double);
    if (f_out) CU_TRY(cudaMemcpy2DAsync(f_out, row, nv.f, pitch, row, B, cudaMemcpyDeviceToHost, c->stream));
    if (nu_out) CU_TRY(cudaMemcpy2DAsync(nu_out, row, nv.bvec, pitch, row, B, cudaMemcpyDeviceToHost, c->stream));
    if (tau_out) CU_TRY(cudaMemcpy2DAsync(tau_out, row, nv.W, pitch, row, B, cudaMemcpyDeviceToHost, c->stream));
    return fetch_results(c, B, c->dOut, 1, nullptr, cubic_ops_out, calc_cov ? 1 : 0, chain_status);
}

// ------------------------------------------------------------------------------------------------
// C ABI: estimators
// ------------------------------------------------------------------------------------------------
// K(theta) and chol(K) -> slot L_K (estimators.py:205-206).  With overlap = true the factorisation is queued on the
// aux stream (the caller must make the main stream wait on ev_lk_done before it reads the slot's L_K).
static int full_front(apm_ctx* c, const double* theta, int B, const int* slots, bool overlap = false) {
    APM_TRY(reset_status(c, B));
    APM_TRY(upload_slots(c, slots, B, c->dSlotsA, false));
    for (int b = 0; b < B; b++) slot_flags(c)[slots[b]] = 0;
    APM_TRY(upload_kernel_params(c, theta, B, c->kind));
    APM_TRY(build_K(c, B, c->kind, c->eps));
    if (overlap) {
        CU_TRY(cudaEventRecord(c->ev_k_ready, c->stream));
        CU_TRY(cudaStreamWaitEvent(c->aux_stream, c->ev_k_ready, 0));
        c->launch_stream = c->aux_stream;
    }
    int rc = run_chol(c, B, c->dK, (long long)c->mat, nullptr, c->dSlotLK, (long long)c->mat, c->dSlotsA, nullptr, 0,
                      c->dSlotLdK, c->dSlotsA, APM_CHAIN_CHOL_K, nullptr);
    if (overlap) {
        c->launch_stream = nullptr;
        if (rc == APM_OK) CU_TRY(cudaEventRecord(c->ev_lk_done, c->aux_stream));
    }
    return rc;
}

// ut_ready: the transposed auxiliary normals of chains 0..B-1 already sit in c->dUT (written there by the native sampler,
// sampler.cuh); u is ignored
static int estimate_full_impl(apm_ctx* c, const double* theta, const double* u, int u_on_device, int N, int B,
                              const int* slots, double* logml_out, int* cubic_ops_out, int* chain_status, bool ut_ready);

extern "C" int apm_estimate_full(apm_ctx* c, const double* theta, const double* u, int u_on_device, int N, int B,
                                 const int* slots, double* logml_out, int* cubic_ops_out, int* chain_status) {
    if (!u) return APM_ERR_INVALID;
    return estimate_full_impl(c, theta, u, u_on_device, N, B, slots, logml_out, cubic_ops_out, chain_status, false);
}

static int estimate_full_impl(apm_ctx* c, const double* theta, const double* u, int u_on_device, int N, int B,
                              const int* slots, double* logml_out, int* cubic_ops_out, int* chain_status, bool ut_ready) {
    APM_TRY(not_companion(c));
    APM_TRY(check_B(c, B));
    if (!theta || !slots || !logml_out) return APM_ERR_INVALID;
    cancel_prefetch(c);
    const bool overlap = c->overlap_chol_k;
    APM_TRY(full_front(c, theta, B, slots, overlap));
    if (N <= 0 || N > c->maxN) {
        set_err("N (importance samples) out of range for this context");
        return APM_ERR_INVALID;
    }
    c->pend_u = (u_on_device || ut_ready) ? nullptr : u;
    c->pend_N = N; c->pend_B = B;
    int rc_mode = (c->approx == 1) ? run_ep(c, B)                               // extension: EP behind post_approx_func
                                   : run_newton(c, B, c->dSlotsA, overlap);     // estimators.py:207 -> lpa.py:81-102
    c->pend_u = nullptr;
    APM_TRY(rc_mode);
    // u is only needed by the importance-sampling tail: a host buffer has had the whole mode search to cross the bus
    if (!ut_ready) APM_TRY(stage_u(c, u, u_on_device, N, B));
    if (c->factored_cov) {
        // chol(C) = L_K U^-T straight from chol(K) and W (lpa.py:111-112 + estimators.py:209 without forming C)
        if (overlap) CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_lk_done, 0));
        const bool hybrid = c->approx == 0 && c->hybrid_newton;
        APM_TRY(run_covariance_factored(c, B, c->dSlotsA, hybrid ? c->dDoneM : nullptr));
    } else {
        APM_TRY(run_covariance(c, B, c->dSlotLC, (long long)c->mat, c->dSlotsA));    // lpa.py:111-112
        // chol(C) in place in the slot                                                estimators.py:209
        APM_TRY(run_chol(c, B, c->dSlotLC, (long long)c->mat, c->dSlotsA, c->dSlotLC, (long long)c->mat, c->dSlotsA, nullptr,
                         0, c->dSlotLdC, c->dSlotsA, APM_CHAIN_CHOL_C, nullptr));
    }
    dim3 cg((c->np + 255) / 256, B);
    prof_begin(c, KID_MISC);
    k_copy_vec<<<cg, 256, 0, c->stream>>>(c->dVec[V_F], c->np, nullptr, c->dSlotMu, c->np, c->dSlotsA, c->np, c->dStatus);
    APM_TRY(check_launch(c, "k_copy_vec"));
    if (overlap) CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_lk_done, 0));
    APM_TRY(run_is_tail(c, N, B, c->dSlotsA, c->dOut, nullptr, 0, c->factored_cov));
    std::vector<int> st(B);
    APM_TRY(fetch_results(c, B, c->dOut, 1, logml_out, cubic_ops_out, 3, st.data()));  // iters + 1 + 2 (est.py:217)
    for (int b = 0; b < B; b++) {
        slot_modes(c)[slots[b]] = c->factored_cov ? 1 : 0;
        slot_flags(c)[slots[b]] = (st[b] == 0);
        if (chain_status) chain_status[b] = st[b];
    }
    return APM_OK;
}

static int cached_common(apm_ctx* c, const int* slots, const double* u, int u_on_device, int N, int B, double* logml_out,
                         double* logw_out, int* chain_status, bool ut_ready = false) {
    APM_TRY(check_B(c, B));
    if (!slots || (!u && !ut_ready)) return APM_ERR_INVALID;
    if (N <= 0 || N > c->maxN) {
        set_err("N (importance samples) out of range for this context");
        return APM_ERR_INVALID;
    }
    cancel_prefetch(c);
    APM_TRY(reset_status(c, B));
    APM_TRY(upload_slots(c, slots, B, c->dSlotsA, true));
    // all slots factored (written by apm_estimate_full) or all explicit (imported): use them as they are; a mixed batch
    // converts its factored slots to explicit chol(C) first
    int n_fact = 0;
    for (int b = 0; b < B; b++) {
        if (slot_modes(c)[slots[b]] == 2) {
            set_err("slot " + std::to_string(slots[b]) + " holds a prior-MC cache (chol K only): no posterior approximation to sample from");
            return APM_ERR_INVALID;
        }
        n_fact += slot_modes(c)[slots[b]] == 1;
    }
    if (n_fact != 0 && n_fact != B) {
        for (int b = 0; b < B; b++) APM_TRY(slot_make_explicit(c, slots[b]));
        n_fact = 0;
    }
    if (!ut_ready) APM_TRY(stage_u(c, u, u_on_device, N, B));
    APM_TRY(run_is_tail(c, N, B, c->dSlotsA, c->dOut, logw_out ? c->dLogw : nullptr, 0, n_fact == B));
    if (logw_out) CU_TRY(cudaMemcpyAsync(logw_out, c->dLogw, sizeof(double) * (size_t)B * N, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaMemsetAsync(c->dIters, 0, sizeof(int) * B, c->stream));
    return fetch_results(c, B, c->dOut, 1, logml_out, nullptr, 0, chain_status);
}

extern "C" int apm_estimate_cached(apm_ctx* c, const int* slots, const double* u, int u_on_device, int N, int B,
                                   double* logml_out, int* chain_status) {
    if (!logml_out) return APM_ERR_INVALID;
    return cached_common(c, slots, u, u_on_device, N, B, logml_out, nullptr, chain_status);
}

extern "C" int apm_estimate_cached_weights(apm_ctx* c, const int* slots, const double* u, int u_on_device, int N, int B,
                                           double* logw_out) {
    if (!logw_out) return APM_ERR_INVALID;
    return cached_common(c, slots, u, u_on_device, N, B, nullptr, logw_out, nullptr);
}

extern "C" int apm_laplace_lml(apm_ctx* c, const double* theta, int B, double* lml_out, int* cubic_ops_out,
                               int* chain_status) {
    APM_TRY(not_companion(c));
    APM_TRY(check_B(c, B));
    if (!theta || !lml_out) return APM_ERR_INVALID;
    APM_TRY(reset_status(c, B));
    APM_TRY(upload_kernel_params(c, theta, B, c->kind));
    APM_TRY(build_K(c, B, c->kind, c->eps));
    APM_TRY(run_newton(c, B));
    NewtonVecs nv = make_nv(c);
    prof_begin(c, KID_NEWTON_VEC);
    k_laplace_lml<<<B, 256, 0, c->stream>>>(nv, c->dLdB, c->nb, c->nb, c->dOut);
    APM_TRY(check_launch(c, "k_laplace_lml"));
    return fetch_results(c, B, c->dOut, 1, lml_out, cubic_ops_out, 0, chain_status);
}

extern "C" int apm_estimate_prior_mc(apm_ctx* c, const double* theta, const int* slots, const double* u, int u_on_device,
                                     int N, int B, double* logml_out, int* chain_status) {
    APM_TRY(not_companion(c));
    APM_TRY(check_B(c, B));
    if (!slots || !u || !logml_out) return APM_ERR_INVALID;
    cancel_prefetch(c);
    if (theta) {
        APM_TRY(full_front(c, theta, B, slots));
    } else {
        // cached chol(K): the slot must hold a valid cache (any mode: the prior-MC tail only reads its L_K)
        APM_TRY(reset_status(c, B));
        APM_TRY(upload_slots(c, slots, B, c->dSlotsA, true));
    }
    APM_TRY(stage_u(c, u, u_on_device, N, B));
    APM_TRY(run_is_tail(c, N, B, c->dSlotsA, c->dOut, nullptr, 1));
    CU_TRY(cudaMemsetAsync(c->dIters, 0, sizeof(int) * B, c->stream));
    std::vector<int> st(B);
    APM_TRY(fetch_results(c, B, c->dOut, 1, logml_out, nullptr, 0, st.data()));
    for (int b = 0; b < B; b++) {
        if (theta) {
            // the slot now holds chol(K) of this theta only: valid for prior-MC re-use, with no chol(C) / f_post behind it
            slot_modes(c)[slots[b]] = 2;
            slot_flags(c)[slots[b]] = (st[b] == 0);
        }
        if (chain_status) chain_status[b] = st[b];
    }
    return APM_OK;
}

// ------------------------------------------------------------------------------------------------
// C ABI: slots
// ------------------------------------------------------------------------------------------------
__global__ void k_logdet_parts(const double* L, int np, double* parts) {
    const int k = blockIdx.x, r = k * 64 + threadIdx.x;  // 64 threads
    double lg = log(L[(size_t)r * np + r]);
    lg = warp_sum(lg);
    __shared__ double red[2];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lg;
    __syncthreads();
    if (threadIdx.x == 0) parts[k] = red[0] + red[1];
}

extern "C" int apm_slot_export(apm_ctx* c, int slot, double* K_chol, double* C_chol, double* f_post, double* logdets2) {
    APM_TRY(not_companion(c));
    if (!c || slot < 0 || slot >= c->nslots) return APM_ERR_INVALID;
    CU_TRY(cudaSetDevice(c->device));
    const size_t n = c->n;
    dim3 grid((c->n + 255) / 256, c->n);
    double* stage = c->dZ;  // scratch (dense n x n)
    if (K_chol) {
        prof_begin(c, KID_MISC);
        k_export_lower<<<grid, 256, 0, c->stream>>>(c->dSlotLK + (size_t)slot * c->mat, c->np, c->n, stage, 0);
        APM_TRY(check_launch(c, "k_export_lower"));
        CU_TRY(cudaMemcpyAsync(K_chol, stage, sizeof(double) * n * n, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
    }
    if ((C_chol || f_post) && slot_modes(c)[slot] == 2) {
        set_err("slot holds a prior-MC cache (chol K only)");
        return APM_ERR_INVALID;
    }
    if (C_chol) {
        APM_TRY(slot_make_explicit(c, slot));     // a factored slot forms chol(C) = L_K V^-1 only now
        prof_begin(c, KID_MISC);
        k_export_lower<<<grid, 256, 0, c->stream>>>(c->dSlotLC + (size_t)slot * c->mat, c->np, c->n, stage, 0);
        APM_TRY(check_launch(c, "k_export_lower"));
        CU_TRY(cudaMemcpyAsync(C_chol, stage, sizeof(double) * n * n, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
    }
    if (f_post) CU_TRY(cudaMemcpyAsync(f_post, c->dSlotMu + (size_t)slot * c->np, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    if (logdets2) {
        std::vector<double> a(c->nb), b(c->nb);
        CU_TRY(cudaMemcpyAsync(a.data(), c->dSlotLdK + (size_t)slot * c->nb, sizeof(double) * c->nb, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaMemcpyAsync(b.data(), c->dSlotLdC + (size_t)slot * c->nb, sizeof(double) * c->nb, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
        logdets2[0] = logdets2[1] = 0.0;
        for (int k = 0; k < c->nb; k++) {
            logdets2[0] += a[k];
            logdets2[1] += b[k];
        }
    }
    CU_TRY(cudaStreamSynchronize(c->stream));
    return APM_OK;
}

extern "C" int apm_slot_import(apm_ctx* c, int slot, const double* K_chol, const double* C_chol, const double* f_post) {
    APM_TRY(not_companion(c));
    if (!c || slot < 0 || slot >= c->nslots || !K_chol) return APM_ERR_INVALID;
    CU_TRY(cudaSetDevice(c->device));
    APM_TRY(import_matrices(c, K_chol, 0, 1, c->dSlotLK + (size_t)slot * c->mat, (long long)c->mat));
    prof_begin(c, KID_MISC);
    k_logdet_parts<<<c->nb, 64, 0, c->stream>>>(c->dSlotLK + (size_t)slot * c->mat, c->np, c->dSlotLdK + (size_t)slot * c->nb);
    APM_TRY(check_launch(c, "k_logdet_parts"));
    if (C_chol) {
        APM_TRY(import_matrices(c, C_chol, 0, 1, c->dSlotLC + (size_t)slot * c->mat, (long long)c->mat));
        prof_begin(c, KID_MISC);
        k_logdet_parts<<<c->nb, 64, 0, c->stream>>>(c->dSlotLC + (size_t)slot * c->mat, c->np, c->dSlotLdC + (size_t)slot * c->nb);
        APM_TRY(check_launch(c, "k_logdet_parts"));
    }
    if (f_post) {
        CU_TRY(cudaMemsetAsync(c->dSlotMu + (size_t)slot * c->np, 0, sizeof(double) * c->np, c->stream));
        CU_TRY(cudaMemcpyAsync(c->dSlotMu + (size_t)slot * c->np, f_post, sizeof(double) * c->n, cudaMemcpyHostToDevice, c->stream));
    }
    CU_TRY(cudaStreamSynchronize(c->stream));
    slot_modes(c)[slot] = C_chol ? 0 : 2;      // chol(K) alone: a prior-MC cache
    slot_flags(c)[slot] = 1;
    return APM_OK;
}

extern "C" int apm_slot_factor(apm_ctx* c, int slot, const double* K, const double* C, int on_device, const double* f_post,
                               int* chain_status) {
    APM_TRY(not_companion(c));
    if (!c || slot < 0 || slot >= c->nslots || !K || !C || !f_post) return APM_ERR_INVALID;
    CU_TRY(cudaSetDevice(c->device));
    slot_flags(c)[slot] = 0;
    APM_TRY(reset_status(c, 1));
    c->hInts[0] = slot;
    CU_TRY(cudaMemcpyAsync(c->dSlotsA, c->hInts, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    APM_TRY(import_matrices(c, K, on_device, 1, c->dK, (long long)c->mat));
    APM_TRY(import_matrices(c, C, on_device, 1, c->dLB, (long long)c->mat));
    APM_TRY(run_chol(c, 1, c->dK, (long long)c->mat, nullptr, c->dSlotLK, (long long)c->mat, c->dSlotsA, nullptr, 0,
                     c->dSlotLdK, c->dSlotsA, APM_CHAIN_CHOL_K, nullptr));
    APM_TRY(run_chol(c, 1, c->dLB, (long long)c->mat, nullptr, c->dSlotLC, (long long)c->mat, c->dSlotsA, nullptr, 0,
                     c->dSlotLdC, c->dSlotsA, APM_CHAIN_CHOL_C, nullptr));
    CU_TRY(cudaMemsetAsync(c->dSlotMu + (size_t)slot * c->np, 0, sizeof(double) * c->np, c->stream));
    CU_TRY(cudaMemcpyAsync(c->dSlotMu + (size_t)slot * c->np, f_post, sizeof(double) * c->n, cudaMemcpyHostToDevice, c->stream));
    int st = 0;
    CU_TRY(cudaMemcpyAsync(c->hInts, c->dStatus, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    st = c->hInts[0];
    slot_modes(c)[slot] = 0;
    slot_flags(c)[slot] = (st == 0);
    if (chain_status) chain_status[0] = st;
    return APM_OK;
}

extern "C" int apm_slot_copy(apm_ctx* c, const int* src, const int* dst, int B) {
    APM_TRY(not_companion(c));
    if (!c || !src || !dst || B <= 0) return APM_ERR_INVALID;
    CU_TRY(cudaSetDevice(c->device));
    for (int b = 0; b < B; b++) {
        if (src[b] < 0 || src[b] >= c->nslots || dst[b] < 0 || dst[b] >= c->nslots) return APM_ERR_INVALID;
        if (src[b] == dst[b]) continue;
        const size_t s = src[b], d = dst[b];
        CU_TRY(cudaMemcpyAsync(c->dSlotLK + d * c->mat, c->dSlotLK + s * c->mat, sizeof(double) * c->mat, cudaMemcpyDeviceToDevice, c->stream));
        CU_TRY(cudaMemcpyAsync(c->dSlotLC + d * c->mat, c->dSlotLC + s * c->mat, sizeof(double) * c->mat, cudaMemcpyDeviceToDevice, c->stream));
        CU_TRY(cudaMemcpyAsync(c->dSlotMu + d * c->np, c->dSlotMu + s * c->np, sizeof(double) * c->np, cudaMemcpyDeviceToDevice, c->stream));
        CU_TRY(cudaMemcpyAsync(c->dSlotMt + d * c->np, c->dSlotMt + s * c->np, sizeof(double) * c->np, cudaMemcpyDeviceToDevice, c->stream));
        slot_modes(c)[d] = slot_modes(c)[s];
        CU_TRY(cudaMemcpyAsync(c->dSlotLdK + d * c->nb, c->dSlotLdK + s * c->nb, sizeof(double) * c->nb, cudaMemcpyDeviceToDevice, c->stream));
        CU_TRY(cudaMemcpyAsync(c->dSlotLdC + d * c->nb, c->dSlotLdC + s * c->nb, sizeof(double) * c->nb, cudaMemcpyDeviceToDevice, c->stream));
        slot_flags(c)[d] = slot_flags(c)[s];
    }
    return APM_OK;
}

// dev/tuning entry (not part of the reference-facing surface): time `reps` batched Cholesky factorisations
// of the K matrices currently in the context (after apm_kernel_build).  mode >> 4 selects what is factorised:
// 0: chol(K) -> slots 0..B-1, 1: chol(I + Ws K Ws) -> LB with inverse diagonal blocks (a Newton round), 2: as 1 with 1/8 of
// the chains active (a straggler round).
extern "C" int apm_dev_chol_bench(apm_ctx* c, int B, int reps, int mode, double* ms_out) {
    APM_TRY(not_companion(c));
    APM_TRY(check_B(c, B));
    if (B > c->nslots || !ms_out) return APM_ERR_INVALID;
    APM_TRY(reset_status(c, B));
    for (int b = 0; b < B; b++) c->hInts[b] = b;
    CU_TRY(cudaMemcpyAsync(c->dSlotsA, c->hInts, sizeof(int) * B, cudaMemcpyHostToDevice, c->stream));
    cudaEvent_t e0, e1;
    CU_TRY(cudaEventCreate(&e0));
    CU_TRY(cudaEventCreate(&e1));
    int rc = APM_OK;
    const int variant = mode >> 4;
    if (variant >= 1) {
        k_fill_double<<<(unsigned)(((size_t)B * c->np + 255) / 256), 256, 0, c->stream>>>(c->dVec[V_WS], 0.5, (long long)B * c->np);
        k_fill_int<<<(B + 255) / 256, 256, 0, c->stream>>>(c->dActive, 1, B);
        if (variant == 2) {
            std::vector<int> act(B, 0);
            for (int b = 0; b < B; b += 8) act[b] = 1;
            cudaMemcpyAsync(c->dActive, act.data(), sizeof(int) * B, cudaMemcpyHostToDevice, c->stream);
            cudaStreamSynchronize(c->stream);
        }
    }
    for (int r = 0; r < reps + 1 && rc == APM_OK; r++) {
        if (r == 1) cudaEventRecord(e0, c->stream);
        if (variant == 0)
            rc = run_chol(c, B, c->dK, (long long)c->mat, nullptr, c->dSlotLK, (long long)c->mat, c->dSlotsA, nullptr, 0,
                          c->dSlotLdK, c->dSlotsA, APM_CHAIN_CHOL_K, nullptr, nullptr);
        else
            rc = run_chol(c, B, c->dK, (long long)c->mat, nullptr, c->dLB, (long long)c->mat, nullptr, c->dVec[V_WS], 1,
                          c->dLdB, nullptr, APM_CHAIN_CHOL_B, c->dActive, c->dInvB);
    }
    cudaEventRecord(e1, c->stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_out = ms / reps;
    return rc;
}

// dev: time k_syrk_sub (a pure tile GEMM, depth n) at a forced occupancy (extra dynamic smem)
extern "C" int apm_dev_syrk_bench(apm_ctx* c, int B, int reps, int smem_bytes, double* ms_out) {
    APM_TRY(not_companion(c));
    APM_TRY(check_B(c, B));
    APM_TRY(reset_status(c, B));
    CU_TRY(cudaFuncSetAttribute(k_syrk_sub, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SyrkParams s;
    s.S = c->dK; s.s_bs = (long long)c->mat; s.lds = c->np;
    s.Z = c->dK; s.z_bs = (long long)c->mat; s.ldz = c->np;
    s.C = c->dLB; s.c_bs = (long long)c->mat; s.ldc = c->np; s.c_idx = nullptr;
    s.nb = c->nb; s.ntiles = c->nb * (c->nb + 1) / 2;
    s.status = c->dStatus;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int r = 0; r < reps + 1; r++) {
        if (r == 1) cudaEventRecord(e0, c->stream);
        k_syrk_sub<<<B * s.ntiles, TILE_THREADS, smem_bytes, c->stream>>>(s);
    }
    cudaEventRecord(e1, c->stream);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_out = ms / reps;
    return cudaGetLastError() == cudaSuccess ? APM_OK : APM_ERR_CUDA;
}

// dev: DMMA TFLOP/s with `warps_per_sm` resident warps (one CTA per SM) and nacc (1, 4, 16) independent
// accumulator chains per warp
extern "C" int apm_dev_dmma_sweep(int device, int warps_per_sm, int nacc, double* tflops) {
    CU_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    double* d = nullptr;
    CU_TRY(cudaMalloc(&d, 64));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000 * 16 / nacc, blocks = prop.multiProcessorCount, threads = warps_per_sm * 32;
    double best = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        if (nacc == 1) k_peak_dmma_n<1><<<blocks, threads>>>(d, iters);
        else if (nacc == 4) k_peak_dmma_n<4><<<blocks, threads>>>(d, iters);
        else k_peak_dmma_n<16><<<blocks, threads>>>(d, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = (double)blocks * warps_per_sm * iters * nacc * 512.0 / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return cudaGetLastError() == cudaSuccess ? APM_OK : APM_ERR_CUDA;
}


// ------------------------------------------------------------------------------------------------
// fp64 peak probes
// ------------------------------------------------------------------------------------------------
extern "C" int apm_measure_fp64_peak(int device, int kind, double* tflops) {
    if (!tflops) return APM_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_err("no CUDA device visible");
        return APM_ERR_NOGPU;
    }
    CU_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    double* d = nullptr;
    CU_TRY(cudaMalloc(&d, 64));
    cudaEvent_t e0, e1;
    CU_TRY(cudaEventCreate(&e0));
    CU_TRY(cudaEventCreate(&e1));
    const int iters = 20000, threads = 256, blocks = prop.multiProcessorCount * 4;
    double best = 0.0;
    for (int rep = 0; rep < 4; rep++) {
        CU_TRY(cudaEventRecord(e0));
        if (kind == 0) k_peak_dmma<<<blocks, threads>>>(d, iters);
        else k_peak_dfma<<<blocks, threads>>>(d, iters);
        CU_TRY(cudaEventRecord(e1));
        CU_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
        // DMMA: 8 mma per iter per warp, 2*8*8*4 flop each; DFMA: 16 fma per iter per thread, 2 flop each
        const double flop = (kind == 0) ? (double)blocks * (threads / 32) * iters * 8.0 * 512.0
                                        : (double)blocks * threads * iters * 16.0 * 2.0;
        const double tf = flop / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return APM_OK;
}

#include "sampler.cuh"
