// r1_chol_engine.cuh -- DEV ONLY (not part of the product): the round-1 cp.async / mma.sync dataflow Cholesky and the separate
// M' = I + L_K^T W L_K SYRK kernel (k_syrk_lk; the product accumulates M' inside k_chol_flow<true>), kept as the A/B baseline of scripts/dev_chol_flow.cu (profiles/: "old" timings).  The product uses csrc/chol_flow.cuh.
#pragma once
#include "../auxiliary-pm-mcmc_b200/csrc/tile_engine.cuh"
#define APM_SKEL 0

namespace apm {

// acc -= X X^T with X = the 64x64 tile in shared memory (stride TSP, conflict-free fragment loads): the
// contribution of a freshly solved panel block L_ik to its own diagonal block.
// Only the lower triangle of the diagonal block is ever read (potrf64_smem), so the warp that owns the upper-right
// 32x32 quadrant does nothing and the two diagonal warps skip their strictly upper 8x8 tiles: 36 of 64 tiles.
#ifndef APM_SYRK_LOWER
#define APM_SYRK_LOWER 1
#endif
__device__ __forceinline__ void syrk_from_tile(Acc& acc, const double* Ts) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    if (APM_SYRK_LOWER && wm < wn) return;
    const bool diag = APM_SYRK_LOWER && wm == wn;
    const double* a_s = Ts + (wm * 32 + g) * TSP + t;
    const double* b_s = Ts + (wn * 32 + g) * TSP + t;
#pragma unroll 4
    for (int kk = 0; kk < TB / 4; kk++) {
        double a[4], b[4];
#pragma unroll
        for (int mi = 0; mi < 4; mi++) a[mi] = -a_s[mi * 8 * TSP + kk * 4];
#pragma unroll
        for (int ni = 0; ni < 4; ni++) b[ni] = b_s[ni * 8 * TSP + kk * 4];
#pragma unroll
        for (int mi = 0; mi < 4; mi++)
#pragma unroll
            for (int ni = 0; ni < 4; ni++)
                if (mi >= ni || !diag) dmma884(acc.v[mi][ni][0], acc.v[mi][ni][1], a[mi], b[ni]);
    }
}


// ------------------------------------------------------------------------------------------------
// In-shared-memory 64x64 Cholesky and triangular solve, blocked by 8-column panels (runtime loop of 8
// panels -> compact code).  Per panel: a left-looking DMMA update of the 64x8 panel with everything to its
// left, then a thread-per-row step on the 8x8 diagonal sub-block (each thread re-derives the 8x8 factor in
// registers from a shared-memory broadcast, so the only communication is the block barrier).
// ------------------------------------------------------------------------------------------------
struct PotrfScratch {
    int fail;
};

// Cholesky factor of an 8x8 SPD block held in d[i][j] (i >= j used); returns L in l[][] and 1/L_jj in inv[].
// L_jj = d * rsqrt(d) (<= 1.5 ulp from sqrt), non-positive / NaN pivots reported through `bad`.
__device__ __forceinline__ void chol8_regs(const double (&d)[8][8], double (&l)[8][8], double (&inv)[8], bool& bad) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
        double piv = d[j][j];
#pragma unroll
        for (int k = 0; k < j; k++) piv = fma(-l[j][k], l[j][k], piv);
        if (!(piv > 0.0)) bad = true;
        const double r = rsqrt(piv);
        inv[j] = r;
        l[j][j] = piv * r;
#pragma unroll
        for (int i = j + 1; i < 8; i++) {
            double v = d[i][j];
#pragma unroll
            for (int k = 0; k < j; k++) v = fma(-l[i][k], l[j][k], v);
            l[i][j] = v * r;
        }
    }
}

// Ts (64x64, stride TSP) <- chol(Ts) in place; strict upper triangle zeroed.  All 128 threads call this, after
// a __syncthreads() that made Ts complete.  Right-looking over 8-column panels with the tile register-resident
// (warp w owns rows 16w..16w+15 as 2 x 8 DMMA accumulator tiles): per panel the 8 columns go through shared
// memory, every row's lane re-derives the 8x8 diagonal factor from a broadcast read and solves its row
// (chol8_regs), the solved panel is published in Lp (64 x 8, stride 12: conflict-free fragment loads) and
// subtracted from all panels to its right by DMMA (independent accumulators, chains of length 2).
// Two block barriers per panel.  Lp: scratch of 64*12 doubles.
constexpr int LPS = 12;
__device__ __forceinline__ void potrf64_smem(double* Ts, double* Lp, PotrfScratch* sc) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int r = warp * 16 + lane;  // row owned by lanes 0..15 in the solve step
    for (int p = 0; p < 8; p++) {
        const int c0 = p * 8;
        double x[8];
        bool bad = false;
        const bool mine = lane < 16 && r >= c0;
        if (mine) {
            double d[8][8], l[8][8], inv[8];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j <= i; j++) d[i][j] = Ts[(c0 + i) * TSP + c0 + j];
            chol8_regs(d, l, inv, bad);
            const double* row = Ts + r * TSP + c0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                double v = row[j];
#pragma unroll
                for (int k = 0; k < j; k++) v = fma(-x[k], l[j][k], v);
                x[j] = v * inv[j];
            }
#pragma unroll
            for (int j = 0; j < 8; j++) Lp[r * LPS + j] = x[j];
            if (bad && r == c0) sc->fail = 1;
        }
        __syncthreads();   // Lp published; every reader of the 8x8 block is done
        if (lane < 16) {
            double* row = Ts + r * TSP + c0;
#pragma unroll
            for (int j = 0; j < 8; j++) row[j] = (mine && c0 + j <= r) ? x[j] : 0.0;
        }
        // trailing update of the lower-triangular tiles to the right: T[mt rows][q cols] -= Lp[mt rows] Lp[q rows]^T
        // (shared-memory resident: 2 DMMAs per 8x8 tile, all tiles independent)
        if (p < 7) {
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
                const int r0 = warp * 16 + mt * 8;
                const double a0 = -Lp[(r0 + g) * LPS + t], a1 = -Lp[(r0 + g) * LPS + 4 + t];
                for (int q = p + 1; q * 8 <= r0; q++) {
                    double* dp = Ts + (r0 + g) * TSP + q * 8 + 2 * t;
                    double2 v = *reinterpret_cast<double2*>(dp);
                    dmma884(v.x, v.y, a0, Lp[(q * 8 + g) * LPS + t]);
                    dmma884(v.x, v.y, a1, Lp[(q * 8 + g) * LPS + 4 + t]);
                    *reinterpret_cast<double2*>(dp) = v;
                }
            }
        }
        __syncthreads();
    }
}


// pack a lower-triangular tile held in shared memory (stride TSP) + reciprocal diagonal
__device__ __forceinline__ void pack_tile(double* Lpk, double* invd, const double* Ts) {
    for (int e = threadIdx.x; e < TB * TB / 2; e += TILE_THREADS) {
        const int r = e >> 5, c = (e & 31) * 2;
        const int qb = r >> 3, pb = c >> 3;
        if (pb <= qb)
            *reinterpret_cast<double2*>(Lpk + lpk_block(qb, pb) + (r & 7) * 8 + (c & 7)) =
                *reinterpret_cast<const double2*>(Ts + r * TSP + c);
    }
    if (threadIdx.x < TB) invd[threadIdx.x] = 1.0 / Ts[threadIdx.x * TSP + threadIdx.x];
}


// ------------------------------------------------------------------------------------------------
// Blocked left-looking Cholesky, launch `k` of nb (k = -1 .. nb-2):
//   CTAs (chain b, row block i = k+1 .. nb-1):
//     if k >= 0 :  L_ik = (A_ik - sum_{j<k} L_ij L_kj^T) L_kk^{-T}
//     if i==k+1 :  L_ii = chol(A_ii - sum_{j<=k} L_ij L_ij^T)          (look-ahead for the next launch)
//   A = diag(scale) * src * diag(scale) (+ I)   -- so B = I + W^1/2 K W^1/2 (lpa.py:91) is never stored.
// ------------------------------------------------------------------------------------------------
struct CholParams {
    const double* src; long long src_bs; int lds; const int* src_idx;
    double* dst; long long dst_bs; int ldd; const int* dst_idx;
    const double* scale; long long scale_bs;   // W^1/2 per chain (null: none)
    int add_identity;
    int nb;
    double* logdet_parts; int logdet_stride; const int* logdet_idx;   // [chain or slot][nb] partial sums of log L_jj
    double* inv_out; long long inv_bs;          // optional: (L_kk^{-1})^T of every diagonal block, [chain][nb][64*64]
    int* status; int fail_code;                 // per-chain status (skip chain if non-zero)
    const int* active;                          // optional Newton mask (skip chain if 0)
    int nchains;
    // per-SM "GEMM token" semaphore (null: off): at most sem_limit of an SM's co-resident CTAs run their panel GEMM at
    // the same time, which staggers the phases of equally long tasks (see gemm_token_acquire)
    int* sm_sem; int sem_limit;
};

// Co-resident CTAs of this kernel work on equally long tasks and, once started together, stay in lock-step: all of
// them in the DMMA-bound panel GEMM (sharing the pipe three ways), then all of them in the latency-bound staging /
// triangular-solve phases (pipe idle).  A counting semaphore per SM (global memory, indexed by %smid) lets only
// sem_limit CTAs into the GEMM phase at once, so the others run their solve phases beside it.  Token holders never
// wait on anything, so there is no deadlock; acquire/release are one thread + the block barriers that exist anyway.
__device__ __forceinline__ unsigned smid() {
    unsigned v;
    asm volatile("mov.u32 %0, %%smid;\n" : "=r"(v));
    return v;
}
__device__ __forceinline__ void gemm_token_acquire(int* sem, int limit) {
    if (threadIdx.x == 0) {
        int* s = sem + smid();
        while (atomicAdd(s, 1) >= limit) {
            atomicSub(s, 1);
            __nanosleep(400);
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void gemm_token_release(int* sem) {   // call after a block barrier that ends the GEMM
    if (threadIdx.x == 0) atomicSub(sem + smid(), 1);
}

// Dependency tracking of the single-launch ("dataflow") variant: progress[chain][row] = number of finished
// column blocks of that block row; a task spins (one thread, acquire loads) until its operands exist.
struct CholFlow {
    int* counter;        // task queue head (zeroed before the launch)
    int* progress;       // [nchains][nb], zeroed before the launch
    const int* skip;     // [nchains] snapshot taken before the launch: non-zero -> chain is not factorised
    int group;           // chains per scheduling group (group-major, step-major inside a group)
    int total_tasks;
    int flags;           // tuning switches (dev): 1 = load the source tile before waiting, 2 = no L2 prefetch
    int spin_ns;
};

// relaxed polling load (no L1 invalidation per poll); the acquire fence is issued once after the spin
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
// all threads call; returns when *p >= need
__device__ __forceinline__ void wait_progress(const int* p, int need, int spin_ns = 100) {
    if (threadIdx.x == 0) {
        if (ld_relaxed_gpu(p) < need) {
            do { __nanosleep(spin_ns); } while (ld_relaxed_gpu(p) < need);
        }
        __threadfence();
    }
    __syncthreads();
}
// two counters, one fence and one barrier
__device__ __forceinline__ void wait_progress2(const int* p0, int need0, const int* p1, int need1, int spin_ns = 100) {
    if (threadIdx.x == 0) {
        while (ld_relaxed_gpu(p0) < need0) __nanosleep(spin_ns);
        while (ld_relaxed_gpu(p1) < need1) __nanosleep(spin_ns);
        __threadfence();
    }
    __syncthreads();
}
// all threads call after their global stores: the block barrier orders every thread's stores before thread 0's
// gpu-scope release store (cumulativity), so one fence-carrying store publishes the whole tile
__device__ __forceinline__ void publish_progress(int* p, int v) {
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(p, v);
}

#ifdef APM_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[16];
__device__ unsigned long long g_phase_counts[16];
#define PHASE_MARK(id)                                                          \
    do {                                                                        \
        __syncthreads();                                                        \
        if (threadIdx.x == 0) {                                                 \
            const long long now__ = clock64();                                  \
            atomicAdd(&g_phase_cycles[id], (unsigned long long)(now__ - t_phase)); \
            atomicAdd(&g_phase_counts[id], 1ull);                               \
            t_phase = now__;                                                    \
        }                                                                       \
    } while (0)
#else
#define PHASE_MARK(id) do {} while (0)
#endif

// One task of the blocked Cholesky: block L_ik of chain b (k >= 0), plus the diagonal block L_ii if i == k+1.
template <bool FLOW>
__device__ __forceinline__ void chol_task(const CholParams& p, const CholFlow& f, int k, int b, int i, double* smem) {
    const double* src = p.src + chain_index(p.src_idx, b) * p.src_bs;
    double* dst = p.dst + chain_index(p.dst_idx, b) * p.dst_bs;
    const double* sc = p.scale ? p.scale + (long long)b * p.scale_bs : nullptr;
    int* prog = FLOW ? f.progress + (size_t)b * p.nb : nullptr;
    TileScratch s = carve_scratch(smem);
    Acc acc;
#ifdef APM_PHASE_TIMING
    long long t_phase = clock64();
#endif
    // The running diagonal block D_ii = A_ii - sum_{j<=k} L_ij L_ij^T lives in dst(i,i): every panel task adds
    // its own contribution right after solving L_ik, so the task that finishes a row (i == k+1) only has the
    // 64x64 factorisation left (no long diagonal GEMM on the critical path).
    const double* sci = sc ? sc + i * TB : nullptr;
    double* dii = dst + (size_t)i * TB * p.ldd + i * TB;
    if (k >= 0) {
        const bool early = FLOW && (f.flags & 1);
        // the source tile never depends on other tasks (in-place: nobody has written tile (i,k) yet)
        if (early) acc_load_tile(acc, src + (size_t)i * TB * p.lds + k * TB, p.lds, sci, sc ? sc + k * TB : nullptr, false);
        if (FLOW) {
            // block row k complete (including L_kk); our own row up to column block k-1 (and its D_ii updates)
            wait_progress2(prog + k, k + 1, prog + i, k, f.spin_ns);
        }
        if (!early) acc_load_tile(acc, src + (size_t)i * TB * p.lds + k * TB, p.lds, sci, sc ? sc + k * TB : nullptr, false);
        PHASE_MARK(0);  // waits + source tile load issue
        prefetch_tile_l2(dst + (size_t)k * TB * p.ldd + k * TB, p.ldd);
        if (k == 0) prefetch_tile_l2(src + (size_t)i * TB * p.lds + i * TB, p.lds);
        else prefetch_tile_l2(dii, p.ldd);
        const bool token = p.sm_sem != nullptr && k > 0;
        if (token) gemm_token_acquire(p.sm_sem, p.sem_limit);
        gemm_nt_64x64<true>(acc, dst + (size_t)i * TB * p.ldd, p.ldd, dst + (size_t)k * TB * p.ldd, p.ldd, k * TB, smem);
        if (token) gemm_token_release(p.sm_sem);
        PHASE_MARK(1);  // panel GEMM
        tile_put_acc(s.Ts, acc);
#if !(APM_SKEL & 8)
        load_diag_block(s.LT, s.invd, dst + (size_t)k * TB * p.ldd + k * TB, p.ldd);
#endif
        __syncthreads();
        PHASE_MARK(2);  // stage T and L_kk
#if !(APM_SKEL & 1)
        trsm64_smem(s.Ts, s.LT, s.invd);
#endif
        __syncthreads();
        PHASE_MARK(3);  // triangular solve
        if (k == 0) acc_load_tile(acc, src + (size_t)i * TB * p.lds + i * TB, p.lds, sci, sci, p.add_identity != 0);
        else acc_load_tile(acc, dii, p.ldd, nullptr, nullptr, false);
        tile_store(s.Ts, dst + (size_t)i * TB * p.ldd + k * TB, p.ldd);
#if !(APM_SKEL & 2)
        syrk_from_tile(acc, s.Ts);
#endif
        PHASE_MARK(4);  // store L_ik + diagonal contribution
        if (i != k + 1) {
            acc_store_tile(acc, dii, p.ldd);
            if (FLOW) publish_progress(prog + i, k + 1);
            PHASE_MARK(5);
        }
    } else {
        acc_load_tile(acc, src + (size_t)i * TB * p.lds + i * TB, p.lds, sci, sci, p.add_identity != 0);
    }
    if (i == k + 1) {
        __syncthreads();   // every warp is done reading Ts as the SYRK operand
        tile_put_acc(s.Ts, acc);
        PotrfScratch* potrf_sc = reinterpret_cast<PotrfScratch*>(s.invd + TB);
        if (threadIdx.x == 0) potrf_sc->fail = 0;
        __syncthreads();
#if !(APM_SKEL & 4)
        potrf64_smem(s.Ts, s.LT, potrf_sc);
#endif
        PHASE_MARK(6);  // 64x64 Cholesky
        tile_store(s.Ts, dii, p.ldd);
        if (FLOW) publish_progress(prog + i, i + 1);
        PHASE_MARK(7);  // store + publish
        if (threadIdx.x < 64) {
            double lg = log(s.Ts[threadIdx.x * TSP + threadIdx.x]);
            lg = warp_sum(lg);
            if ((threadIdx.x & 31) == 0) s.invd[threadIdx.x >> 5] = lg;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (p.logdet_parts)
                p.logdet_parts[(size_t)chain_index(p.logdet_idx, b) * p.logdet_stride + i] = s.invd[0] + s.invd[1];
            if (potrf_sc->fail) atomicMax(&p.status[b], p.fail_code);
        }
        if (p.inv_out && !(APM_SKEL & 4)) {
            // (L_ii^{-1})^T = I * L_ii^{-T}: lets the single right-hand-side solves of the Newton step
            // (k_trsv2) replace 64-step substitutions on the diagonal blocks by parallel 64x64 mat-vecs
            __syncthreads();
            pack_tile(s.LT, s.invd, s.Ts);
            __syncthreads();
            for (int e = threadIdx.x; e < TB * TB; e += TILE_THREADS) {
                const int r = e >> 6, c = e & 63;
                s.Ts[r * TSP + c] = (r == c) ? 1.0 : 0.0;
            }
            __syncthreads();
            trsm64_smem(s.Ts, s.LT, s.invd);
            __syncthreads();
            tile_store(s.Ts, p.inv_out + (long long)b * p.inv_bs + (size_t)i * TB * TB, TB);
        }
        PHASE_MARK(8);  // log-det + inverse of the diagonal block
    }
    __syncthreads();  // scratch is re-used by the next task of a persistent CTA
}

// one launch per block column (k = -1 .. nb-2); the heavy CTAs (look-ahead diagonal) come first in launch order
__global__ void __launch_bounds__(TILE_THREADS, MIN_CTAS) k_chol_step(CholParams p, int k) {
    extern __shared__ __align__(16) double smem[];
    const int rows_per_chain = p.nb - k - 1;
    int b, i;
    if ((int)blockIdx.x < p.nchains) {
        b = blockIdx.x;
        i = k + 1;
    } else {
        const int r = blockIdx.x - p.nchains;
        b = r / (rows_per_chain - 1);
        i = k + 2 + r % (rows_per_chain - 1);
    }
    if (p.status[b] != 0) return;
    if (p.active && !p.active[b]) return;
    CholFlow f = {};
    chol_task<false>(p, f, k, b, i, smem);
}

// The whole batched factorisation as ONE cooperative launch: persistent CTAs pull tasks from a queue ordered
// so that every dependency has a lower index (group-major; inside a group step-major with the look-ahead
// diagonals first), and wait on per-row progress counters instead of kernel boundaries.  No launch tails,
// the diagonal critical path starts as early as its operands exist, and a group's matrices stay L2-resident.
__global__ void __launch_bounds__(TILE_THREADS, MIN_CTAS) k_chol_dataflow(CholParams p, CholFlow f) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_task;
    const int nb = p.nb, G = f.group;
    const int per_group = G * (1 + nb * (nb - 1) / 2);
    for (;;) {
        if (threadIdx.x == 0) s_task = atomicAdd(f.counter, 1);
        __syncthreads();
        const int t = s_task;
        __syncthreads();
        if (t >= f.total_tasks) return;
        const int grp = t / per_group;
        int r = t - grp * per_group;
        int k, b, i;
        if (r < G) {
            k = -1; i = 0; b = grp * G + r;
        } else {
            r -= G;
            k = 0;
            while (r >= G * (nb - k - 1)) { r -= G * (nb - k - 1); k++; }
            if (r < G) {
                b = grp * G + r; i = k + 1;
            } else {
                r -= G;
                b = grp * G + r / (nb - k - 2);
                i = k + 2 + r % (nb - k - 2);
            }
        }
        if (b >= p.nchains || f.skip[b]) continue;
        chol_task<true>(p, f, k, b, i, smem);
    }
}

// skip[b] = chain b must not be factorised (failed earlier, or converged in the Newton loop)
__global__ void k_chol_skip_snapshot(const int* status, const int* active, int* skip, int n) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n) skip[b] = (status[b] != 0) || (active && !active[b]);
}


// ------------------------------------------------------------------------------------------------
// M' straight from L_K ("TN" operands): M = I + L_K^T W L_K is a sum of outer products of ROWS of the row-major L_K,
//   M[i][j] = [i == j] + sum_{k >= max(i,j)} W_k L_K[k][i] L_K[k][j],
// so the k-chunks are staged k-major ([KC][64] per operand, row stride 68: the fragment reads a = S[k = t][m = g] are two-way
// = minimal for 8-byte accesses) and W_k scales the A fragment.  No transposed, scaled copy Y' of L_K (k_make_Y) and no
// second read of it.  The CTA of lower tile (i', j') of M' = P M P computes tile (I, J) = (nb-1-i', nb-1-j') of M (I <= J: its
// k-range starts at block J) and stores it index-reversed.
// ------------------------------------------------------------------------------------------------
constexpr int TNS = 68;
constexpr int TN_STAGE_DOUBLES = 2 * KC * TNS + KC;   // A chunk, B chunk, W chunk
static_assert(STAGES * TN_STAGE_DOUBLES * 8 <= TILE_SMEM_BYTES, "k-major stages exceed the shared memory of a CTA");

__device__ __forceinline__ void gemm_tn_load_stage(double* st, const double* __restrict__ A, const double* __restrict__ Bm, int ld,
                                                   const double* __restrict__ w, int k0, int tid) {
    // 16 rows x 32 16-byte segments per operand = 512 segments, 4 per thread
#pragma unroll
    for (int q = 0; q < (KC * 32) / TILE_THREADS; q++) {
        const int seg = tid + q * TILE_THREADS;
        const int r = seg >> 5, c = (seg & 31) * 2;
        cp_async16(st + r * TNS + c, A + (size_t)(k0 + r) * ld + c);
        cp_async16(st + KC * TNS + r * TNS + c, Bm + (size_t)(k0 + r) * ld + c);
    }
    if (tid < KC / 2) cp_async16(st + 2 * KC * TNS + tid * 2, w + k0 + tid * 2);
}

// acc += sum_{k < kdepth} w[k] * A[k][0..63] (x) Bm[k][0..63]   (A, Bm: pointers to row k = 0 of the two column blocks)
__device__ __forceinline__ void gemm_tn_64x64_w(Acc& acc, const double* __restrict__ A, const double* __restrict__ Bm, int ld,
                                                const double* __restrict__ w, int kdepth, double* smem) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    const int nchunks = kdepth / KC;
    if (nchunks == 0) return;
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nchunks) gemm_tn_load_stage(smem + s * TN_STAGE_DOUBLES, A, Bm, ld, w, s * KC, tid);
        cp_async_commit();
    }
    for (int kc = 0; kc < nchunks; kc++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nk = kc + STAGES - 1;
            if (nk < nchunks) gemm_tn_load_stage(smem + (nk % STAGES) * TN_STAGE_DOUBLES, A, Bm, ld, w, nk * KC, tid);
            cp_async_commit();
        }
        const double* st = smem + (kc % STAGES) * TN_STAGE_DOUBLES;
        const double* a_s = st + t * TNS + wm * 32 + g;
        const double* b_s = st + KC * TNS + t * TNS + wn * 32 + g;
        const double* w_s = st + 2 * KC * TNS + t;
#pragma unroll
        for (int kk = 0; kk < KC / 4; kk++) {
            const double wk = w_s[kk * 4];
            double a[4], b[4];
#pragma unroll
            for (int mi = 0; mi < 4; mi++) a[mi] = a_s[kk * 4 * TNS + mi * 8] * wk;
#pragma unroll
            for (int ni = 0; ni < 4; ni++) b[ni] = b_s[kk * 4 * TNS + ni * 8];
#pragma unroll
            for (int mi = 0; mi < 4; mi++)
#pragma unroll
                for (int ni = 0; ni < 4; ni++) dmma884(acc.v[mi][ni][0], acc.v[mi][ni][1], a[mi], b[ni]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();
}

struct SyrkLkParams {
    const double* LK; long long lk_bs; int ldk; const int* lk_idx;   // chol(K) (slot)
    const double* W; long long w_bs;                                  // Newton / EP weights, [chain][np] (0 in the padding)
    double* M; long long m_bs; int ldm;                               // M' (lower tiles, full diagonal tiles)
    int nb; int ntiles;
    const int* status; const int* mask;
};

__global__ void __launch_bounds__(TILE_THREADS, MIN_CTAS) k_syrk_lk(SyrkLkParams p) {
    extern __shared__ __align__(16) double smem[];
    const int b = blockIdx.x / p.ntiles;
    const int tix = p.ntiles - 1 - (int)(blockIdx.x % p.ntiles);   // deepest tiles first
    if (p.status[b] != 0 || (p.mask && !p.mask[b])) return;
    int i = (int)((sqrt(8.0 * tix + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= tix) i++;
    while (i * (i + 1) / 2 > tix) i--;
    const int j = tix - i * (i + 1) / 2;                           // lower tile (i, j) of M'
    const int I = p.nb - 1 - i, J = p.nb - 1 - j;                  // tile (I, J) of M, I <= J
    const double* L = p.LK + chain_index(p.lk_idx, b) * p.lk_bs;
    const double* W = p.W + (long long)b * p.w_bs;
    double* M = p.M + (long long)b * p.m_bs;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3, wm = warp >> 1, wn = warp & 1;
    Acc acc;
    acc.zero();
    if (i == j && wm == wn) {
#pragma unroll
        for (int mi = 0; mi < 4; mi++) {   // identity on the diagonal of the 32x32 warp block
            if (g == 2 * t) acc.v[mi][mi][0] = 1.0;
            if (g == 2 * t + 1) acc.v[mi][mi][1] = 1.0;
        }
    }
    const int kstart = J * TB;
    gemm_tn_64x64_w(acc, L + (size_t)kstart * p.ldk + I * TB, L + (size_t)kstart * p.ldk + J * TB, p.ldk, W + kstart,
                    (p.nb - J) * TB, smem);
    // M'[i*64 + 63 - m][j*64 + 63 - n] = M[I*64 + m][J*64 + n]
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            const int m = wm * 32 + mi * 8 + g, n = wn * 32 + ni * 8 + 2 * t;
            *reinterpret_cast<double2*>(M + (size_t)(i * TB + 63 - m) * p.ldm + j * TB + 62 - n) =
                make_double2(acc.v[mi][ni][1], acc.v[mi][ni][0]);
        }
}


}  // namespace apm
