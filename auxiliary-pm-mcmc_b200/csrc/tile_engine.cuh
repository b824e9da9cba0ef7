// tile_engine.cuh -- blocked fp64 dense kernels on 64x64 tiles, batched over chains.
//
// Everything O(n^3) in the reference's hot path (LAPACK dpotrf/dpotrs and BLAS dgemm reached through
// scipy: lpa.py:92, 111-112; estimators.py:206, 209, 223, 225) is expressed with four kernels that share
// one DMMA micro-kernel (acc += A * B^T on a 64x64 tile, both operands row-major with k contiguous).  The Cholesky
// factorisations live in chol_flow.cuh (TMA / mbarrier dataflow kernel); this file holds the remaining tile kernels:
//
//   k_trsm_rows   X L^T = R for a 64-row panel of right-hand sides, one CTA walks all block columns
//   k_syrk_sub    C = S - Z Z^T (lower tiles)
//   k_trsm_rev    L_C = L_K V^-1 (explicit chol(C) of a factored cache, on demand)
//   k_gemm_tri    F = mu + U^T L^T (lower-triangular L, block column ranges clipped)
#pragma once
#include "common.cuh"

namespace apm {

// ------------------------------------------------------------------------------------------------
// DMMA micro-kernel: acc(64x64) += A(64 x kdepth) * B(64 x kdepth)^T.  128 threads = 2x2 warps, each
// warp owns a 32x32 block = 4x4 m8n8 tiles.  A/B k-chunks are staged by a 4-deep cp.async pipeline.
// ------------------------------------------------------------------------------------------------
struct Acc {
    double v[4][4][2];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) v[i][j][0] = v[i][j][1] = 0.0;
    }
};

__device__ __forceinline__ void gemm_load_stage(double* smA, double* smB, const double* __restrict__ A, int lda,
                                                const double* __restrict__ Bm, int ldb, int k0, int tid) {
    constexpr int SEG_PER_ROW = KC / 2;                     // 16-byte segments per row of a k-chunk
    constexpr int NSEG = TB * SEG_PER_ROW;                  // per operand per stage
#pragma unroll
    for (int q = 0; q < NSEG / TILE_THREADS; q++) {
        const int seg = tid + q * TILE_THREADS;
        const int row = seg / SEG_PER_ROW, sg = seg % SEG_PER_ROW, cs = sg * 2;
        const int ds = SWIZZLE ? ((sg ^ ((row & 3) << 1)) * 2) : cs;
        cp_async16(smA + row * KCP + ds, A + (size_t)row * lda + k0 + cs);
        cp_async16(smB + row * KCP + ds, Bm + (size_t)row * ldb + k0 + cs);
    }
}

template <bool NEG = false>
__device__ __forceinline__ void gemm_nt_64x64(Acc& acc, const double* __restrict__ A, int lda,
                                              const double* __restrict__ Bm, int ldb, int kdepth, double* smem) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    double* smA = smem;
    double* smB = smem + STAGES * TB * KCP;
    const int nchunks = kdepth / KC;
    if (nchunks == 0) return;
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nchunks) gemm_load_stage(smA + s * TB * KCP, smB + s * TB * KCP, A, lda, Bm, ldb, s * KC, tid);
        cp_async_commit();
    }
    for (int kc = 0; kc < nchunks; kc++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nk = kc + STAGES - 1;
            if (nk < nchunks) {
                const int st = nk % STAGES;
                gemm_load_stage(smA + st * TB * KCP, smB + st * TB * KCP, A, lda, Bm, ldb, nk * KC, tid);
            }
            cp_async_commit();
        }
        const int st = kc % STAGES;
        const double* a_s = smA + st * TB * KCP + (wm * 32 + g) * KCP + (SWIZZLE ? (t & 1) : t);
        const double* b_s = smB + st * TB * KCP + (wn * 32 + g) * KCP + (SWIZZLE ? (t & 1) : t);
        const int swz = (g & 3) << 1, th = t >> 1;
#pragma unroll
        for (int kk = 0; kk < KC / 4; kk++) {
            double a[4], b[4];
            const int ko = SWIZZLE ? (((kk * 2 + th) ^ swz) * 2) : kk * 4;
#pragma unroll
            for (int mi = 0; mi < 4; mi++) a[mi] = NEG ? -a_s[mi * 8 * KCP + ko] : a_s[mi * 8 * KCP + ko];
#pragma unroll
            for (int ni = 0; ni < 4; ni++) b[ni] = b_s[ni * 8 * KCP + ko];
#pragma unroll
            for (int mi = 0; mi < 4; mi++)
#pragma unroll
                for (int ni = 0; ni < 4; ni++) dmma884(acc.v[mi][ni][0], acc.v[mi][ni][1], a[mi], b[ni]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();  // all warps done with the stage buffers: they may be re-used by the caller
}

// ------------------------------------------------------------------------------------------------
// 64x64 work tile in shared memory (row stride TSP) -- coalesced global I/O and fragment exchange
// ------------------------------------------------------------------------------------------------
// Ts = diag(rs) * S * diag(cs) (+ I on a diagonal tile).  rs / cs may be null.
__device__ __forceinline__ void tile_load(double* Ts, const double* __restrict__ S, int ld, const double* rs,
                                          const double* cs, bool add_identity) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = lane * 2;
    const double c0 = cs ? cs[c] : 1.0, c1 = cs ? cs[c + 1] : 1.0;
#pragma unroll 4
    for (int rr = 0; rr < 16; rr++) {
        const int r = warp * 16 + rr;
        double2 v = *reinterpret_cast<const double2*>(S + (size_t)r * ld + c);
        if (rs || cs) {
            const double rsv = rs ? rs[r] : 1.0;
            v.x = rsv * v.x * c0;
            v.y = rsv * v.y * c1;
        }
        if (add_identity) {
            if (r == c) v.x += 1.0;
            if (r == c + 1) v.y += 1.0;
        }
        Ts[r * TSP + c] = v.x;
        Ts[r * TSP + c + 1] = v.y;
    }
}
__device__ __forceinline__ void tile_fill_rowvec(double* Ts, const double* vec) {  // Ts[r][c] = vec[c]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double v0 = vec[lane * 2], v1 = vec[lane * 2 + 1];
#pragma unroll 4
    for (int rr = 0; rr < 16; rr++) {
        const int r = warp * 16 + rr;
        Ts[r * TSP + lane * 2] = v0;
        Ts[r * TSP + lane * 2 + 1] = v1;
    }
}
__device__ __forceinline__ void tile_store(const double* Ts, double* __restrict__ Dm, int ld) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 4
    for (int rr = 0; rr < 16; rr++) {
        const int r = warp * 16 + rr;
        double2 v;
        v.x = Ts[r * TSP + lane * 2];
        v.y = Ts[r * TSP + lane * 2 + 1];
        *reinterpret_cast<double2*>(Dm + (size_t)r * ld + lane * 2) = v;
    }
}
// acc <- diag(rs) * S * diag(cs) (+ I on a diagonal tile), loaded straight from global memory in the DMMA
// accumulator layout (each lane: 16-byte pieces C[g][2t..2t+1]).  Issued BEFORE the k-loop so the DRAM
// latency hides behind the GEMM; the k-loop then subtracts (gemm with NEG=true): acc = S - A B^T.
__device__ __forceinline__ void acc_load_tile(Acc& acc, const double* __restrict__ S, int ld, const double* rs,
                                              const double* cs, bool add_identity) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
        const int r = wm * 32 + mi * 8 + g;
        const double rsv = rs ? rs[r] : 1.0;
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            const int c = wn * 32 + ni * 8 + 2 * t;
            double2 v = __ldcg(reinterpret_cast<const double2*>(S + (size_t)r * ld + c));
            if (rs || cs) {
                v.x = rsv * v.x * (cs ? cs[c] : 1.0);
                v.y = rsv * v.y * (cs ? cs[c + 1] : 1.0);
            }
            if (add_identity) {
                if (r == c) v.x += 1.0;
                if (r == c + 1) v.y += 1.0;
            }
            acc.v[mi][ni][0] = v.x;
            acc.v[mi][ni][1] = v.y;
        }
    }
}
// acc <- vec[c] broadcast along rows (F = mu + ...)
__device__ __forceinline__ void acc_load_rowvec(Acc& acc, const double* vec) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = lane & 3, wn = warp & 1;
#pragma unroll
    for (int ni = 0; ni < 4; ni++) {
        const double2 v = vec ? *reinterpret_cast<const double2*>(vec + wn * 32 + ni * 8 + 2 * t) : make_double2(0.0, 0.0);
#pragma unroll
        for (int mi = 0; mi < 4; mi++) {
            acc.v[mi][ni][0] = v.x;
            acc.v[mi][ni][1] = v.y;
        }
    }
}
// Ts <- acc
__device__ __forceinline__ void tile_put_acc(double* Ts, const Acc& acc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++)
            *reinterpret_cast<double2*>(Ts + (wm * 32 + mi * 8 + g) * TSP + wn * 32 + ni * 8 + 2 * t) =
                make_double2(acc.v[mi][ni][0], acc.v[mi][ni][1]);
}
// acc -> global, directly from the fragment layout (16-byte stores, 64-byte row segments per quad)
__device__ __forceinline__ void acc_store_tile(const Acc& acc, double* __restrict__ Dm, int ld) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++)
            *reinterpret_cast<double2*>(Dm + (size_t)(wm * 32 + mi * 8 + g) * ld + wn * 32 + ni * 8 + 2 * t) =
                make_double2(acc.v[mi][ni][0], acc.v[mi][ni][1]);
}
// warm L2 with a 64x64 tile that will be needed after the k-loop (2 lines of 128 B per thread)
__device__ __forceinline__ void prefetch_tile_l2(const double* __restrict__ S, int ld) {
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int e = threadIdx.x + q * TILE_THREADS;      // 256 lines: row = e / 4, 128-byte piece = e % 4
        const double* ptr = S + (size_t)(e >> 2) * ld + (e & 3) * 16;
        asm volatile("prefetch.global.L2 [%0];\n" ::"l"(ptr));
    }
}

// Ts += sign * acc (fragment layout of the DMMA accumulators)
__device__ __forceinline__ void tile_add_acc(double* Ts, const Acc& acc, double sign) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            double* p = Ts + (wm * 32 + mi * 8 + g) * TSP + wn * 32 + ni * 8 + 2 * t;
            p[0] += sign * acc.v[mi][ni][0];
            p[1] += sign * acc.v[mi][ni][1];
        }
}

// Ts <- Ts * L^{-T} for a 64x64 tile; Ld = L (64x64 lower, row-major, stride TSP), invd[c] = 1 / L[c][c].
// Right-looking over 8-column panels with the whole tile register-resident: every warp owns 16 rows
// (2 m-tiles x 8 n-tiles of DMMA accumulators) and needs nothing from the other warps, so the only
// synchronisation is __syncwarp.  Per panel: the 16x8 panel goes through shared memory to 16 lanes that solve
// it exactly against the 8x8 diagonal sub-block (substitution, one row per lane), comes back as DMMA
// A-fragments, and is subtracted from all panels to its right (independent accumulators, chains of length 2).
// The caller must __syncthreads() before (Ts, Ld complete) and after (Ts readable by other warps).
// The 64x64 lower-triangular operand L is held PACKED in shared memory: its 36 lower 8x8 blocks, block (q,p)
// (q >= p) at index q(q+1)/2 + p, 64 doubles each, row-major (18 KB instead of a padded 34 KB tile).
__device__ __forceinline__ int lpk_block(int q, int p) { return (q * (q + 1) / 2 + p) * 64; }
constexpr int LPK_DOUBLES = 36 * 64;

__device__ __forceinline__ void trsm64_smem(double* Ts, const double* Lpk, const double* invd) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    double* rows = Ts + warp * 16 * TSP;
    double acc[2][8][2];
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            const double2 v = *reinterpret_cast<const double2*>(rows + (mt * 8 + g) * TSP + nt * 8 + 2 * t);
            acc[mt][nt][0] = v.x;
            acc[mt][nt][1] = v.y;
        }
#pragma unroll
    for (int p = 0; p < 8; p++) {
        const int c0 = p * 8;
        if (p > 0) {
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
                *reinterpret_cast<double2*>(rows + (mt * 8 + g) * TSP + c0 + 2 * t) = make_double2(acc[mt][p][0], acc[mt][p][1]);
            __syncwarp();
        }
        if (lane < 16) {
            double* row = rows + lane * TSP + c0;
            const double* l8 = Lpk + lpk_block(p, p);   // broadcast reads
            double x[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                double v = row[j];
#pragma unroll
                for (int k = 0; k < j; k++) v = fma(-x[k], l8[j * 8 + k], v);
                x[j] = v * invd[c0 + j];
            }
#pragma unroll
            for (int j = 0; j < 8; j++) row[j] = x[j];
        }
        __syncwarp();
        if (p < 7) {
            double a[2][2];
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int kk = 0; kk < 2; kk++) a[mt][kk] = -rows[(mt * 8 + g) * TSP + c0 + kk * 4 + t];
#pragma unroll
            for (int q = p + 1; q < 8; q++) {
                const double* blk = Lpk + lpk_block(q, p) + g * 8 + t;
#pragma unroll
                for (int kk = 0; kk < 2; kk++) {
                    const double bq = blk[kk * 4];
#pragma unroll
                    for (int mt = 0; mt < 2; mt++) dmma884(acc[mt][q][0], acc[mt][q][1], a[mt][kk], bq);
                }
            }
        }
    }
}

// stage the diagonal block L_kk (row-major in global memory, ld) into the packed layout + reciprocal diagonal
__device__ __forceinline__ void load_diag_block(double* Lpk, double* invd, const double* __restrict__ Lkk, int ld) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pb = lane >> 2, cc = (lane & 3) * 2;          // 8-column block and column inside it
    double2 v[16];
#pragma unroll
    for (int rr = 0; rr < 16; rr++)   // all 16 loads in flight before the first use
        v[rr] = __ldcg(reinterpret_cast<const double2*>(Lkk + (size_t)(warp * 16 + rr) * ld + lane * 2));
#pragma unroll
    for (int rr = 0; rr < 16; rr++) {
        const int r = warp * 16 + rr, qb = r >> 3;
        if (pb <= qb) *reinterpret_cast<double2*>(Lpk + lpk_block(qb, pb) + (r & 7) * 8 + cc) = v[rr];
        if (r == lane * 2) invd[r] = 1.0 / v[rr].x;
        if (r == lane * 2 + 1) invd[r] = 1.0 / v[rr].y;
    }
}
// carve the epilogue scratch out of the (finished) GEMM stage buffers
struct TileScratch {
    double* Ts;
    double* LT;
    double* invd;
};
__device__ __forceinline__ TileScratch carve_scratch(double* smem) {
    TileScratch s;
    s.Ts = smem;                          // 64 x TSP work tile
    s.LT = smem + TB * TSP;               // packed diagonal block L_kk (LPK_DOUBLES); potrf panel scratch
    s.invd = s.LT + LPK_DOUBLES;
    return s;
}
static_assert((TB * TSP + LPK_DOUBLES + TB) * 8 <= TILE_SMEM_BYTES, "tile scratch exceeds GEMM smem");

// ------------------------------------------------------------------------------------------------
// X L^T = R for 64-row panels: CTA (chain b, row block r) walks block columns k = 0..nb-1:
//   X_rk = (R_rk - sum_{j<k} X_rj L_kj^T) L_kk^{-T}
// R_rk = Rsrc tile scaled by column vector cs (R = K diag(W^1/2), lpa.py:90/111) or a plain tile.
// ------------------------------------------------------------------------------------------------
struct TrsmParams {
    const double* R; long long r_bs; int ldr; const int* r_idx;
    const double* cs; long long cs_bs;          // optional column scaling of R
    double* X; long long x_bs; int ldx;
    const double* L; long long l_bs; int ldl; const int* l_idx;
    int nb;         // block columns (n_pad / 64)
    int row_blocks; // row blocks of R / X per chain
    const int* status; const int* active;
};

__global__ void __launch_bounds__(TILE_THREADS, MIN_CTAS) k_trsm_rows(TrsmParams p) {
    extern __shared__ __align__(16) double smem[];
    const int b = blockIdx.x / p.row_blocks, rb = blockIdx.x % p.row_blocks;
    if (p.status[b] != 0) return;
    if (p.active && !p.active[b]) return;
    const double* R = p.R + chain_index(p.r_idx, b) * p.r_bs + (size_t)rb * TB * p.ldr;
    double* X = p.X + (long long)b * p.x_bs + (size_t)rb * TB * p.ldx;
    const double* L = p.L + chain_index(p.l_idx, b) * p.l_bs;
    const double* cs = p.cs ? p.cs + (long long)b * p.cs_bs : nullptr;
    TileScratch s = carve_scratch(smem);
    Acc acc;
    for (int k = 0; k < p.nb; k++) {
        acc_load_tile(acc, R + k * TB, p.ldr, nullptr, cs ? cs + k * TB : nullptr, false);
        prefetch_tile_l2(L + (size_t)k * TB * p.ldl + k * TB, p.ldl);
        gemm_nt_64x64<true>(acc, X, p.ldx, L + (size_t)k * TB * p.ldl, p.ldl, k * TB, smem);
        tile_put_acc(s.Ts, acc);
        load_diag_block(s.LT, s.invd, L + (size_t)k * TB * p.ldl + k * TB, p.ldl);
        __syncthreads();
        trsm64_smem(s.Ts, s.LT, s.invd);
        __syncthreads();
        tile_store(s.Ts, X + k * TB, p.ldx);
        __threadfence();
        __syncthreads();  // X_rk is an operand of the following block columns
    }
}

// ------------------------------------------------------------------------------------------------
// C = S - Z Z^T, lower tiles (i >= j) only.  (lpa.py:112 with Z = (B^{-1/2} W^{1/2} K)^T)
// ------------------------------------------------------------------------------------------------
struct SyrkParams {
    const double* S; long long s_bs; int lds;
    const double* Z; long long z_bs; int ldz;
    double* C; long long c_bs; int ldc; const int* c_idx;
    int nb; int ntiles;  // nb*(nb+1)/2
    const int* status;
};

__global__ void __launch_bounds__(TILE_THREADS, MIN_CTAS) k_syrk_sub(SyrkParams p) {
    extern __shared__ __align__(16) double smem[];
    const int b = blockIdx.x / p.ntiles;
    int tix = blockIdx.x % p.ntiles;
    if (p.status[b] != 0) return;
    // tix -> (i, j), j <= i : row i starts at i(i+1)/2
    int i = (int)((sqrt(8.0 * tix + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= tix) i++;
    while (i * (i + 1) / 2 > tix) i--;
    const int j = tix - i * (i + 1) / 2;
    const double* S = p.S + (long long)b * p.s_bs;
    const double* Z = p.Z + (long long)b * p.z_bs;
    double* C = p.C + chain_index(p.c_idx, b) * p.c_bs;
    Acc acc;
    acc_load_tile(acc, S + (size_t)i * TB * p.lds + j * TB, p.lds, nullptr, nullptr, false);
    gemm_nt_64x64<true>(acc, Z + (size_t)i * TB * p.ldz, p.ldz, Z + (size_t)j * TB * p.ldz, p.ldz, p.nb * TB, smem);
    acc_store_tile(acc, C + (size_t)i * TB * p.ldc + j * TB, p.ldc);
}

// ------------------------------------------------------------------------------------------------
// Factored posterior covariance (replaces lpa.py:111-112 + estimators.py:209 in the fused FULL estimate):
//   C = (K^-1 + W)^-1 = L_K M^-1 L_K^T,   M = I + L_K^T W L_K        (Woodbury; W of the last Newton step)
//   M = U U^T with U upper triangular  =>  chol(C) = L_C = L_K U^-T   (lower x lower, positive diagonal: unique)
// U comes from an ordinary lower Cholesky of the index-reversed matrix M' = P M P = L' L'^T (U = P L' P), and
//   (L_C P) L'^T = L_K P   is a right triangular solve in reversed column order.
// Cost n^3/3 (M') + n^3/3 (chol) + n^3/3 (solve) instead of n^3 (Z) + n^3 (C) + n^3/3 (chol C); no explicit C,
// no cancellation in K - Z Z^T, and M' has eigenvalues >= 1.
// ------------------------------------------------------------------------------------------------

// (L_C P) L'^T = L_K P : CTA (chain b, row block rb) walks the reversed block columns k' = nb-1-rb .. nb-1.
//   X[rb,k'] = (R[rb,k'] - sum_{j'=nb-1-rb}^{k'-1} X[rb,j'] L'[k',j']^T) L'_{k'k'}^{-T},  R[r][c'] = L_K[r][n-1-c'] (lower part)
// X (reversed coordinates) is kept in a scratch matrix as the GEMM operand and written, columns re-reversed, to L_C.
struct TrsmRevParams {
    const double* LK; long long lk_bs; int ldk; const int* lk_idx;   // L_K (slot)
    const double* Lp; long long lp_bs; int ldp;                       // L' = chol(M')
    double* X; long long x_bs; int ldx;                               // scratch, reversed coordinates
    double* LC; long long lc_bs; int ldc; const int* lc_idx;          // output L_C (slot)
    int nb;
    const int* status;
};

__global__ void __launch_bounds__(TILE_THREADS, MIN_CTAS) k_trsm_rev(TrsmRevParams p) {
    extern __shared__ __align__(16) double smem[];
    const int b = blockIdx.x / p.nb;
    const int rb = p.nb - 1 - (int)(blockIdx.x % p.nb);   // longest rows first
    if (p.status[b] != 0) return;
    const double* LK = p.LK + chain_index(p.lk_idx, b) * p.lk_bs + (size_t)rb * TB * p.ldk;
    const double* Lp = p.Lp + (long long)b * p.lp_bs;
    double* X = p.X + (long long)b * p.x_bs + (size_t)rb * TB * p.ldx;
    double* LC = p.LC + chain_index(p.lc_idx, b) * p.lc_bs + (size_t)rb * TB * p.ldc;
    TileScratch s = carve_scratch(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3, wm = warp >> 1, wn = warp & 1;
    const int k0 = p.nb - 1 - rb;
    Acc acc;
    for (int k = k0; k < p.nb; k++) {
        const int ob = p.nb - 1 - k;   // original column block of L_K / L_C
        // R tile in accumulator layout: R[r][c'] = L_K[r][ob*64 + 63 - c'], zero above L_K's diagonal (ob == rb)
#pragma unroll
        for (int mi = 0; mi < 4; mi++) {
            const int r = wm * 32 + mi * 8 + g;
#pragma unroll
            for (int ni = 0; ni < 4; ni++) {
                const int c = wn * 32 + ni * 8 + 2 * t;          // reversed columns c, c+1 <- original 63-c, 62-c
                const double2 v = __ldcg(reinterpret_cast<const double2*>(LK + (size_t)r * p.ldk + ob * TB + 62 - c));
                const bool diag = (ob == rb);
                acc.v[mi][ni][0] = (!diag || 63 - c <= r) ? v.y : 0.0;
                acc.v[mi][ni][1] = (!diag || 62 - c <= r) ? v.x : 0.0;
            }
        }
        prefetch_tile_l2(Lp + (size_t)k * TB * p.ldp + k * TB, p.ldp);
        gemm_nt_64x64<true>(acc, X + k0 * TB, p.ldx, Lp + (size_t)k * TB * p.ldp + k0 * TB, p.ldp, (k - k0) * TB, smem);
        tile_put_acc(s.Ts, acc);
        load_diag_block(s.LT, s.invd, Lp + (size_t)k * TB * p.ldp + k * TB, p.ldp);
        __syncthreads();
        trsm64_smem(s.Ts, s.LT, s.invd);
        __syncthreads();
        tile_store(s.Ts, X + k * TB, p.ldx);
        {   // L_C[r][ob*64 + c] = X[r][63 - c]
#pragma unroll 4
            for (int rr = 0; rr < 16; rr++) {
                const int r = warp * 16 + rr;
                double2 v;
                v.x = s.Ts[r * TSP + 63 - lane * 2];
                v.y = s.Ts[r * TSP + 62 - lane * 2];
                *reinterpret_cast<double2*>(LC + (size_t)r * p.ldc + ob * TB + lane * 2) = v;
            }
        }
        __threadfence();
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// F[s][i] = mu[i] + sum_{j <= i} U^T[s][j] L[i][j]   (estimators.py:223, 323), tile (row block rb of
// samples, block column k): depth clipped to the lower triangle ((k+1)*64 columns).
// ------------------------------------------------------------------------------------------------
struct GemmTriParams {
    const double* UT; long long u_bs; int ldu;           // [chain][Npad][n_pad]
    const double* L; long long l_bs; int ldl; const int* l_idx;
    const double* mu; long long mu_bs; const int* mu_idx; // null -> 0 (prior MC)
    double* F; long long f_bs; int ldf;
    int nb; int row_blocks;
    const int* status;
};

__global__ void __launch_bounds__(TILE_THREADS, MIN_CTAS) k_gemm_tri(GemmTriParams p) {
    extern __shared__ __align__(16) double smem[];
    const int per_chain = p.nb * p.row_blocks;
    const int b = blockIdx.x / per_chain;
    const int r = blockIdx.x % per_chain;
    const int rb = r / p.nb, k = p.nb - 1 - (r % p.nb);  // deepest tiles first
    if (p.status && p.status[b] != 0) return;
    const double* UT = p.UT + (long long)b * p.u_bs + (size_t)rb * TB * p.ldu;
    const double* L = p.L + chain_index(p.l_idx, b) * p.l_bs + (size_t)k * TB * p.ldl;
    double* F = p.F + (long long)b * p.f_bs + (size_t)rb * TB * p.ldf + k * TB;
    Acc acc;
    acc_load_rowvec(acc, p.mu ? p.mu + chain_index(p.mu_idx, b) * p.mu_bs + k * TB : nullptr);
    gemm_nt_64x64<false>(acc, UT, p.ldu, L, p.ldl, (k + 1) * TB, smem);
    acc_store_tile(acc, F, p.ldf);
}

}  // namespace apm
