// vec_kernels.cuh -- the HBM-bound part of the path: K(theta) build, Newton vector updates, matvecs,
// single right-hand-side triangular solves, the importance-sampling epilogue, layout helpers.
#pragma once
#include <math_constants.h>
#include <cooperative_groups.h>
#include "common.cuh"

namespace apm {

constexpr double HALF_LOG_2PI = 0.91893853320467274178;

// ------------------------------------------------------------------------------------------------
// K(theta): gpdemo/kernels.pyx:12-49 (isotropic) and :52-90 (ARD), batched over chains.
// Arithmetic follows the scalar reference loop exactly (k ascending, division by the length-scale,
// no FMA contraction); sigma = exp(theta0), tau_k = exp(theta_k) and 2*tau^2 are computed on the host
// with libm and passed in kp[chain][*].  Lower tiles are computed and mirrored through shared memory.
// Rows/cols >= n (padding up to a multiple of 64) are set to the identity.
// ------------------------------------------------------------------------------------------------
struct KBuildParams {
    const double* X; int n, D, np, nb;          // X [np][D] (pad rows zero)
    const double* kp; int kp_stride;            // per chain: ARD [sigma, tau_1..tau_D, 1/tau_1..1/tau_D], ISO [sigma, 2 tau^2, 1/(2 tau^2)]
    int ard; double eps;
    double* K; long long k_bs;                  // [chain][np][np]
    int ntiles;                                 // nb(nb+1)/2
};

// a / b with r = RN(1/b) from the host: q = RN(a r), exact remainder by FMA, one correction (Markstein).  Returns the
// correctly rounded quotient (the operands here are far from over/underflow), 3 fp64 operations instead of the ~15 of
// the IEEE division sequence.
__device__ __forceinline__ double div_by(double a, double b, double r) {
    const double q = __dmul_rn(a, r);
    const double rem = fma(-q, b, a);
    return fma(rem, r, q);
}

__global__ void __launch_bounds__(256) k_build_K(KBuildParams p) {
    extern __shared__ __align__(16) double smem[];
    const int b = blockIdx.x / p.ntiles;
    const int tix = blockIdx.x % p.ntiles;
    int ti = (int)((sqrt(8.0 * tix + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= tix) ti++;
    while (ti * (ti + 1) / 2 > tix) ti--;
    const int tj = tix - ti * (ti + 1) / 2;
    const int D = p.D;
    double* Xi = smem;                    // [64][D]
    double* XjT = Xi + 64 * D;            // [D][64]
    double* Ts = XjT + 64 * D;            // [64][65] mirror staging
    double* prm = Ts + 64 * VSP;          // [2D+1]
    const int tid = threadIdx.x;
    // The 64 rows of X a tile needs are contiguous in the row-major, row-padded X: two bulk copies (cp.async.bulk,
    // SASS UBLKCP.S.G) onto an mbarrier bring them in -- X_i where it is used, X_j through the (still unused) mirror staging
    // area, from where it is transposed so that the feature loop reads it conflict-free.
    __shared__ __align__(8) unsigned long long xbar;
    const uint32_t xbar_u = smem_u32(&xbar);
    const uint32_t xbytes = (uint32_t)(64 * D * sizeof(double));
    if (tid == 0) {
        mbar_init(xbar_u, 1);
        fence_mbar_init();
        mbar_expect_tx(xbar_u, 2 * xbytes);
        bulk_load_1d(smem_u32(Xi), p.X + (size_t)ti * 64 * D, xbytes, xbar_u);
        bulk_load_1d(smem_u32(Ts), p.X + (size_t)tj * 64 * D, xbytes, xbar_u);
    }
    const double* kp = p.kp + (size_t)b * p.kp_stride;
    for (int e = tid; e < 2 * D + 1; e += 256) prm[e] = (p.ard || e < 3) ? kp[e] : 0.0;
    __syncthreads();                       // the barrier is initialised (and prm written) before anybody waits on it
    mbar_wait(xbar_u, 0);
    for (int e = tid; e < 64 * D; e += 256) {
        const int r = e / D, k = e % D;
        XjT[k * 64 + r] = Ts[e];
    }
    __syncthreads();
    const double sigma = prm[0];
    const int tx = tid & 31, ty = tid >> 5;
    double* Kb = p.K + (long long)b * p.k_bs;
    // every thread owns 8 rows x 2 columns; the feature loop is outermost so that the length-scale, its reciprocal and
    // the two column coordinates are read once per feature (12 shared-memory reads per 16 elements instead of 64).
    // Per element the terms are still accumulated in ascending k, as the reference's scalar loop does.
    double acc[8][2];
#pragma unroll
    for (int rr = 0; rr < 8; rr++) acc[rr][0] = acc[rr][1] = 0.0;
    const double* xi = Xi + (ty * 8) * D;
    if (p.ard) {
        for (int k = 0; k < D; k++) {
            const double tau = prm[k + 1], rtau = prm[D + 1 + k];
            const double xj0 = XjT[k * 64 + tx * 2], xj1 = XjT[k * 64 + tx * 2 + 1];
#pragma unroll
            for (int rr = 0; rr < 8; rr++) {
                const double x = xi[rr * D + k];
                const double d0 = div_by(__dsub_rn(x, xj0), tau, rtau), d1 = div_by(__dsub_rn(x, xj1), tau, rtau);
                acc[rr][0] = __dadd_rn(acc[rr][0], __dmul_rn(d0, d0));
                acc[rr][1] = __dadd_rn(acc[rr][1], __dmul_rn(d1, d1));
            }
        }
    } else {
        for (int k = 0; k < D; k++) {
            const double xj0 = XjT[k * 64 + tx * 2], xj1 = XjT[k * 64 + tx * 2 + 1];
#pragma unroll
            for (int rr = 0; rr < 8; rr++) {
                const double x = xi[rr * D + k];
                const double d0 = __dsub_rn(x, xj0), d1 = __dsub_rn(x, xj1);
                acc[rr][0] = __dadd_rn(acc[rr][0], __dmul_rn(d0, d0));
                acc[rr][1] = __dadd_rn(acc[rr][1], __dmul_rn(d1, d1));
            }
        }
    }
#pragma unroll
    for (int rr = 0; rr < 8; rr++) {
        const int r = ty * 8 + rr;
        const int gi = ti * 64 + r;
        double out[2];
#pragma unroll
        for (int cc = 0; cc < 2; cc++) {
            const int gj = tj * 64 + tx * 2 + cc;
            double v = p.ard ? __dmul_rn(sigma, exp(-acc[rr][cc] * 0.5)) : __dmul_rn(sigma, exp(div_by(-acc[rr][cc], prm[1], prm[2])));
            if (gi == gj) v = sigma + p.eps;
            if (gi >= p.n || gj >= p.n) v = (gi == gj) ? 1.0 : 0.0;
            out[cc] = v;
        }
        *reinterpret_cast<double2*>(Kb + (size_t)gi * p.np + tj * 64 + tx * 2) = make_double2(out[0], out[1]);
        Ts[r * VSP + tx * 2] = out[0];
        Ts[r * VSP + tx * 2 + 1] = out[1];
    }
    if (ti != tj) {
        __syncthreads();
        // mirrored tile: K[tj*64 + c][ti*64 + r] = Ts[r][c]
#pragma unroll
        for (int cc = 0; cc < 8; cc++) {
            const int c = ty * 8 + cc;
            double2 v;
            v.x = Ts[(tx * 2) * VSP + c];
            v.y = Ts[(tx * 2 + 1) * VSP + c];
            *reinterpret_cast<double2*>(Kb + (size_t)(tj * 64 + c) * p.np + ti * 64 + tx * 2) = v;
        }
    }
}

// dK/dtheta_p for p = 0..P-1 (extension named by the north-star; the reference has no gradients -- SURVEY App. D):
//   ARD:  dK_ij/dtheta_0 = K_ij (off-diagonal), sigma (diagonal);  dK_ij/dtheta_{k+1} = K_ij ((x_ik - x_jk)/tau_k)^2
//   ISO:  dK_ij/dtheta_0 as above;                                 dK_ij/dtheta_1 = K_ij |x_i - x_j|^2 / tau^2
// with K_ij computed exactly as k_build_K does.  Output: dense [chain][P][n][n] (ld = n), one 64x64 tile per CTA.
struct KGradParams {
    const double* X; int n, D, nb;
    const double* kp; int kp_stride;
    int ard; int P;
    double* dK;                                 // [chain][P][n][n]
};

__global__ void __launch_bounds__(256) k_build_dK(KGradParams p) {
    extern __shared__ __align__(16) double smem[];
    const int ntiles = p.nb * p.nb;
    const int b = blockIdx.x / ntiles;
    const int ti = (blockIdx.x % ntiles) / p.nb, tj = (blockIdx.x % ntiles) % p.nb;
    const int D = p.D;
    double* Xi = smem;                    // [64][D]
    double* XjT = Xi + 64 * D;            // [D][64]
    double* prm = XjT + 64 * D;           // [2D+1]
    const int tid = threadIdx.x;
    for (int e = tid; e < 64 * D; e += 256) {
        const int r = e / D, k = e % D;
        Xi[e] = p.X[(size_t)(ti * 64 + r) * D + k];
        XjT[k * 64 + r] = p.X[(size_t)(tj * 64 + r) * D + k];
    }
    const double* kp = p.kp + (size_t)b * p.kp_stride;
    for (int e = tid; e < 2 * D + 1; e += 256) prm[e] = (p.ard || e < 3) ? kp[e] : 0.0;
    __syncthreads();
    const double sigma = prm[0];
    const int tx = tid & 31, ty = tid >> 5;
    const size_t n = p.n;
    double* out = p.dK + (size_t)b * p.P * n * n;
    for (int rr = 0; rr < 8; rr++) {
        const int r = ty * 8 + rr, gi = ti * 64 + r;
        for (int cc = 0; cc < 2; cc++) {
            const int c = tx * 2 + cc, gj = tj * 64 + c;
            if (gi >= p.n || gj >= p.n) continue;
            double acc = 0.0;
            for (int k = 0; k < D; k++) {
                const double diff = __dsub_rn(Xi[r * D + k], XjT[k * 64 + c]);
                const double d = p.ard ? div_by(diff, prm[k + 1], prm[D + 1 + k]) : diff;
                acc = __dadd_rn(acc, __dmul_rn(d, d));
            }
            const double kij = (gi == gj) ? sigma
                                          : (p.ard ? __dmul_rn(sigma, exp(-acc * 0.5)) : __dmul_rn(sigma, exp(div_by(-acc, prm[1], prm[2]))));
            out[(size_t)gi * n + gj] = kij;                                        // d/dtheta_0
            if (p.ard) {
                for (int k = 0; k < D; k++) {
                    const double d = div_by(__dsub_rn(Xi[r * D + k], XjT[k * 64 + c]), prm[k + 1], prm[D + 1 + k]);
                    out[(size_t)(k + 1) * n * n + (size_t)gi * n + gj] = (gi == gj) ? 0.0 : kij * (d * d);
                }
            } else {
                // prm[1] = 2 tau^2:  |x_i - x_j|^2 / tau^2 = 2 acc / (2 tau^2)
                out[n * n + (size_t)gi * n + gj] = (gi == gj) ? 0.0 : kij * (2.0 * acc * prm[2]);
            }
        }
    }
}

// u [chain][n][N] (reference layout, estimators.py:155-160) -> uT [chain][Npad][np], zero padded
__global__ void k_transpose_u(const double* __restrict__ u, long long u_bs, int n, int N, double* __restrict__ uT,
                              long long ut_bs, int np, int Npad) {
    __shared__ double tile[32][33];
    const int b = blockIdx.z;
    const int i0 = blockIdx.x * 32, s0 = blockIdx.y * 32;
    const double* ub = u + (long long)b * u_bs;
    double* utb = uT + (long long)b * ut_bs;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, s = s0 + tx;
        tile[r][tx] = (i < n && s < N) ? ub[(size_t)i * N + s] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int s = s0 + r, i = i0 + tx;
        if (s < Npad && i < np) utb[(size_t)s * np + i] = tile[tx][r];
    }
}

// per-slot partial log-dets of L_C = L_K U^-T:  sum log diag L_C = sum log diag L_K - sum log diag L'
__global__ void k_logdet_combine(const double* ldK, const double* ldM, double* ldC, const int* slot_idx, int nb, const int* status) {
    const int b = blockIdx.x, k = threadIdx.x;
    if (status[b] != 0 || k >= nb) return;
    const long long sl = chain_index(slot_idx, b);
    ldC[sl * nb + k] = ldK[sl * nb + k] - ldM[(long long)b * nb + k];
}

// ------------------------------------------------------------------------------------------------
// Newton iteration pieces (lpa.py:85-99), one CTA per chain for the O(n) parts
// ------------------------------------------------------------------------------------------------
struct NewtonVecs {
    double* f; double* W; double* Ws; double* bvec; double* a; double* t; double* s; double* fnew;
    long long vs;   // stride between chains (= np)
    const double* y;
    int n, np;
    // n_active[0]: still-active chains, [1]: of those, predicted to finish in the next iteration, [2]: chains that
    // finished in this round in the B-space form
    int* active; int* iters; int* status; int* n_active;
    double tol; int max_iters;
    // hybrid Newton (run_newton): every chain runs an iteration in the reference's B-space form or, when the iteration
    // is predicted to be its last, in "M-space"; mask_b / mask_m = the form of chain b's current (then next) iteration,
    // done_m[b] = the chain's last iteration was an M-space one (all null: not hybrid)
    int* done_m; int* mask_b; int* mask_m;
    double pred_factor;
};

// v = exp(-f^2/2 - log_ndtr(y f) - log(2 pi)/2); grad = v y; W = v^2 + grad f; Ws = sqrt(W); b = W f + grad
// rhs_c != 0: also t = c = b / W^1/2, the right-hand side of the mat-vec-free Newton step (see k_fnew_from_s): with
// W^1/2 K W^1/2 = B - I the reference's t = W^1/2 K b (lpa.py:94) equals (B - I) c, hence s = B^-1 t = c - B^-1 c.
__global__ void k_newton_prep(NewtonVecs nv, int rhs_c) {
    const int b = blockIdx.x;
    if (!nv.active[b] || nv.status[b] != 0) return;
    const long long o = (long long)b * nv.vs;
    for (int i = threadIdx.x; i < nv.np; i += blockDim.x) {
        double W = 0.0, Ws = 0.0, bv = 0.0;
        if (i < nv.n) {
            const double f = nv.f[o + i], y = nv.y[i];
            const double v = exp(-0.5 * f * f - log_ndtr(y * f) - HALF_LOG_2PI);
            const double g = v * y;
            W = v * v + g * f;
            Ws = sqrt(W);
            bv = W * f + g;
        }
        nv.W[o + i] = W;
        nv.Ws[o + i] = Ws;
        nv.bvec[o + i] = bv;
        if (rhs_c) nv.t[o + i] = Ws > 0.0 ? bv / Ws : 0.0;
    }
}

// f_new = K a without the mat-vec (lpa.py:95).  With s = B^-1 t, t = W^1/2 K b and a = b - W^1/2 s (lpa.py:94):
//   B s = t  <=>  s + W^1/2 K W^1/2 s = W^1/2 K b  <=>  W^1/2 K a = s,   i.e.   (K a)_i = s_i / W^1/2_i  exactly.
// Dividing amplifies the rounding error of s_i (~ eps cond(B) |s|) by 1 / W^1/2_i, so components whose W^1/2 is below
// `thr` (probit: y_i f_i large and positive, a few per cent of the data) are computed as the row product K[i,:] a instead:
// the result stays within ~1e-13 of the mat-vec while the 8 n^2-byte read of K shrinks to the flagged rows.
// One CTA (256 threads) per chain.
// rhs_c != 0 (mat-vec-free step, see k_newton_prep): k_trsv2 solved with the right-hand side c = b / W^1/2, so nv.s holds
// B^-1 c; s = c - B^-1 c and a = b - W^1/2 s are formed here first (k_trsv2 itself stays as it is: its load schedule is
// sensitive to any change of the kernel).
__global__ void __launch_bounds__(256) k_fnew_from_s(const double* __restrict__ K, long long k_bs, int ld, NewtonVecs nv,
                                                     double thr, int rhs_c) {
    __shared__ int flagged[1024];
    __shared__ int n_flagged;
    const int b = blockIdx.x;
    if (!nv.active[b] || nv.status[b] != 0) return;
    const long long o = (long long)b * nv.vs;
    if (threadIdx.x == 0) n_flagged = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < nv.np; i += 256) {
        double v = 0.0;
        if (rhs_c) {
            const double si = nv.t[o + i] - nv.s[o + i];
            nv.s[o + i] = si;
            nv.a[o + i] = nv.bvec[o + i] - nv.Ws[o + i] * si;
        }
        if (i < nv.n) {
            const double ws = nv.Ws[o + i];
            if (ws >= thr) v = nv.s[o + i] / ws;
            else {
                const int q = atomicAdd(&n_flagged, 1);
                if (q < 1024) flagged[q] = i;
            }
        }
        nv.fnew[o + i] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double* Kb = K + (long long)b * k_bs;
    const double* a = nv.a + o;
    if (n_flagged <= 1024) {
        for (int q = warp; q < n_flagged; q += 8) {
            const int i = flagged[q];
            const double* row = Kb + (size_t)i * ld;
            double acc0 = 0.0, acc1 = 0.0;
            int j = lane * 2;
            for (; j + 1 < nv.n; j += 64) {
                const double2 m = *reinterpret_cast<const double2*>(row + j);
                acc0 = fma(m.x, a[j], acc0);
                acc1 = fma(m.y, a[j + 1], acc1);
            }
            if (j < nv.n) acc0 = fma(row[j], a[j], acc0);
            const double v = warp_sum(acc0 + acc1);
            if (lane == 0) nv.fnew[o + i] = v;
        }
    } else {
        // more small-curvature components than the list holds: every row below the threshold by its own warp
        for (int i = warp; i < nv.n; i += 8) {
            if (nv.Ws[o + i] >= thr) continue;
            const double* row = Kb + (size_t)i * ld;
            double acc = 0.0;
            for (int j = lane; j < nv.n; j += 32) acc = fma(row[j], a[j], acc);
            acc = warp_sum(acc);
            if (lane == 0) nv.fnew[o + i] = acc;
        }
    }
}

// out[i] = rs[i] * sum_j M[i][j] x[j]; grid (np/32, chains), 256 threads (8 warps x 4 rows)
__global__ void __launch_bounds__(256) k_matvec(const double* __restrict__ M, long long m_bs, int ld, int ncols,
                                                const double* __restrict__ x, const double* __restrict__ rs,
                                                double* __restrict__ out, long long vs, const int* active,
                                                const int* status) {
    const int b = blockIdx.y;
    if ((active && !active[b]) || (status && status[b] != 0)) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double* Mb = M + (long long)b * m_bs;
    const double* xb = x + (long long)b * vs;
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const int r = blockIdx.x * 32 + warp * 4 + rr;
        const double* row = Mb + (size_t)r * ld;
        double acc = 0.0;
        for (int j = lane * 2; j < ncols; j += 64) {
            const double2 m = *reinterpret_cast<const double2*>(row + j);
            const double2 xv = *reinterpret_cast<const double2*>(xb + j);
            acc = fma(m.x, xv.x, acc);
            acc = fma(m.y, xv.y, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) out[(long long)b * vs + r] = rs ? rs[(long long)b * vs + r] * acc : acc;
    }
}

// Symmetric mat-vec that reads only the lower tiles of K (half the HBM bytes of k_matvec): CTA (row block i, chain)
// streams the tiles (i, j <= i) once; each tile T gives the direct product T x_j for rows i and, for j < i, the
// transposed product T^T x_i for rows j, which goes to a scratch part[chain][i][j][64] and is added in a fixed order
// by k_symv_reduce (deterministic: no atomics).  256 threads = 8 warps x 8 tile rows.
__global__ void __launch_bounds__(256) k_symv_lower(const double* __restrict__ M, long long m_bs, int ld, int nb,
                                                    const double* __restrict__ x, long long vs, double* __restrict__ direct,
                                                    double* __restrict__ part, const int* active, const int* status) {
    __shared__ double colred[8][64];
    const int b = blockIdx.y;
    if ((active && !active[b]) || (status && status[b] != 0)) return;
    const int i = nb - 1 - blockIdx.x;   // longest rows first
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double* Mb = M + (long long)b * m_bs + (size_t)(i * 64 + warp * 8) * ld;
    const double* xb = x + (long long)b * vs;
    double xi[8];
#pragma unroll
    for (int rr = 0; rr < 8; rr++) xi[rr] = xb[i * 64 + warp * 8 + rr];
    double racc[8];
#pragma unroll
    for (int rr = 0; rr < 8; rr++) racc[rr] = 0.0;
    for (int j = 0; j <= i; j++) {
        const double2 xj = *reinterpret_cast<const double2*>(xb + j * 64 + lane * 2);
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int rr = 0; rr < 8; rr++) {
            const double2 m = *reinterpret_cast<const double2*>(Mb + (size_t)rr * ld + j * 64 + lane * 2);
            racc[rr] = fma(m.x, xj.x, racc[rr]);
            racc[rr] = fma(m.y, xj.y, racc[rr]);
            c0 = fma(m.x, xi[rr], c0);
            c1 = fma(m.y, xi[rr], c1);
        }
        if (j < i) {   // transposed contribution of this tile to rows of block j
            colred[warp][lane * 2] = c0;
            colred[warp][lane * 2 + 1] = c1;
            __syncthreads();
            if (threadIdx.x < 64) {
                double v = 0.0;
#pragma unroll
                for (int w = 0; w < 8; w++) v += colred[w][threadIdx.x];
                part[(((size_t)b * nb + i) * nb + j) * 64 + threadIdx.x] = v;
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int rr = 0; rr < 8; rr++) {
        const double v = warp_sum(racc[rr]);
        if (lane == 0) direct[(long long)b * vs + i * 64 + warp * 8 + rr] = v;
    }
}
// out_j = rs_j * (direct_j + sum_{i > j} part[i][j]); grid (nb, chains), 64 threads
__global__ void k_symv_reduce(const double* __restrict__ direct, const double* __restrict__ part, int nb,
                              const double* __restrict__ rs, double* __restrict__ out, long long vs, const int* active,
                              const int* status) {
    const int b = blockIdx.y, j = blockIdx.x, r = threadIdx.x;
    if ((active && !active[b]) || (status && status[b] != 0)) return;
    double v = direct[(long long)b * vs + j * 64 + r];
    for (int i = j + 1; i < nb; i++) v += part[(((size_t)b * nb + i) * nb + j) * 64 + r];
    out[(long long)b * vs + j * 64 + r] = rs ? rs[(long long)b * vs + j * 64 + r] * v : v;
}

// s = L^{-T} L^{-1} t (lpa.py:94 cho_solve with one right-hand side), then a = b - Ws * s.
// One CTA (256 threads) per chain walks the 64-row blocks of L twice.  The 64x64 diagonal-block solves use the
// explicit (L_kk^{-1})^T blocks written by k_chol_step (B = I + W^1/2 K W^1/2 has eigenvalues >= 1, so its
// diagonal blocks are well conditioned), which turns every step into coalesced, fully parallel mat-vecs.
// dynamic smem: w[np] + rhs[64] + part[32][64] (k_trsv2<false> uses 8 of the 32)
// skip_forward: w = L^-1 t is already in nv.s (forward substitution fused into the factorisation, chol_flow.cuh): only the
// backward half runs (L is read once instead of twice).
template <bool SKIP_FWD>
__global__ void __launch_bounds__(256) k_trsv2(const double* __restrict__ L, long long l_bs, int ld, int nb,
                                               const double* __restrict__ LinvT, long long inv_bs, NewtonVecs nv) {
    extern __shared__ __align__(16) double smem[];
    const int b = blockIdx.x;
    if (!nv.active[b] || nv.status[b] != 0) return;
    const int np = nb * 64;
    double* w = smem;            // [np]
    double* rhs = w + np;        // [64]
    double* part = rhs + 64;     // [8][64]
    const double* Lb = L + (long long)b * l_bs;
    const double* Ib = LinvT + (long long)b * inv_bs;
    const long long o = (long long)b * nv.vs;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < np; i += 256) w[i] = SKIP_FWD ? nv.s[o + i] : nv.t[o + i];
    __syncthreads();
    // ---- forward: L w = t
    if (!SKIP_FWD)
    for (int kb = 0; kb < nb; kb++) {
        // rhs = t_k - L[k, 0:k] w[0:k]; warp handles 8 rows at once (independent loads in flight)
        const int kcols = kb * 64;
        double acc[8];
#pragma unroll
        for (int rr = 0; rr < 8; rr++) acc[rr] = 0.0;
        const double* rows = Lb + (size_t)(kb * 64 + warp * 8) * ld;
        for (int j = lane * 2; j < kcols; j += 64) {
            const double w0 = w[j], w1 = w[j + 1];
#pragma unroll
            for (int rr = 0; rr < 8; rr++) {
                const double2 m = *reinterpret_cast<const double2*>(rows + (size_t)rr * ld + j);
                acc[rr] = fma(m.x, w0, acc[rr]);
                acc[rr] = fma(m.y, w1, acc[rr]);
            }
        }
#pragma unroll
        for (int rr = 0; rr < 8; rr++) {
            const double v = warp_sum(acc[rr]);
            if (lane == 0) rhs[warp * 8 + rr] = w[kb * 64 + warp * 8 + rr] - v;
        }
        __syncthreads();
        // w_k = L_kk^{-1} rhs :  w_c = sum_r LinvT[r][c] rhs[r]  (r <= c); 4 row chunks x 64 columns
        {
            const int c = tid & 63, ch = tid >> 6;
            const double* blk = Ib + (size_t)kb * 4096;
            double a0 = 0.0;
#pragma unroll
            for (int r = ch * 16; r < ch * 16 + 16; r++) a0 = fma(blk[r * 64 + c], rhs[r], a0);
            part[ch * 64 + c] = a0;
        }
        __syncthreads();
        if (tid < 64) w[kb * 64 + tid] = (part[tid] + part[64 + tid]) + (part[128 + tid] + part[192 + tid]);
        __syncthreads();
    }
    // ---- backward: L^T s = w
    for (int kb = nb - 1; kb >= 0; kb--) {
        // rhs = w_k - sum_{i > k} L[i, k]^T s_i : 8 row groups x 32 column pairs, coalesced 512-byte row segments
        {
            const int cg = tid & 31, rg = tid >> 5;
            double a0 = 0.0, a1 = 0.0;
            const int r_end = np;
            const double* col = Lb + kb * 64 + cg * 2;
            if (SKIP_FWD) {
                // backward-only variant: a pure latency chain of 12 column-block reads -- the 8 rows a thread reads of every
                // 64-row block are issued together (the trip count is a multiple of 8 by construction).  Summation order
                // shared with k_trsv_back_c4 (so that the two give the same bits): rows 16 j .. 16 j + 15 of every block go
                // into sub-sum j (what CTA j of the cluster adds up), the 8 row slots are added per j, then (v0 + v1) + (v2 + v3).
                double s0[4] = {0.0, 0.0, 0.0, 0.0}, s1[4] = {0.0, 0.0, 0.0, 0.0};
                for (int ib = kb + 1; ib < nb; ib++) {
                    double2 m[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) m[q] = *reinterpret_cast<const double2*>(col + (size_t)(ib * 64 + q * 8 + rg) * ld);
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const double sr = w[ib * 64 + q * 8 + rg];
                        s0[q >> 1] = fma(m[q].x, sr, s0[q >> 1]);
                        s1[q >> 1] = fma(m[q].y, sr, s1[q >> 1]);
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    part[(j * 8 + rg) * 64 + cg * 2] = s0[j];
                    part[(j * 8 + rg) * 64 + cg * 2 + 1] = s1[j];
                }
            } else {
#pragma unroll 4
                for (int r = (kb + 1) * 64 + rg; r < r_end; r += 8) {
                    const double2 m = *reinterpret_cast<const double2*>(col + (size_t)r * ld);
                    const double sr = w[r];
                    a0 = fma(m.x, sr, a0);
                    a1 = fma(m.y, sr, a1);
                }
            }
            if (!SKIP_FWD) {
                part[rg * 64 + cg * 2] = a0;
                part[rg * 64 + cg * 2 + 1] = a1;
            }
        }
        __syncthreads();
        if (tid < 64) {
            double v;
            if (SKIP_FWD) {
                double vj[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    vj[j] = 0.0;
#pragma unroll
                    for (int q = 0; q < 8; q++) vj[j] += part[(j * 8 + q) * 64 + tid];
                }
                v = (vj[0] + vj[1]) + (vj[2] + vj[3]);
            } else {
                v = 0.0;
#pragma unroll
                for (int q = 0; q < 8; q++) v += part[q * 64 + tid];
            }
            rhs[tid] = w[kb * 64 + tid] - v;
        }
        __syncthreads();
        // s_k = L_kk^{-T} rhs : s_r = sum_c LinvT[r][c] rhs[c]; warp handles 8 rows
        {
            const double* blk = Ib + (size_t)kb * 4096 + (size_t)(warp * 8) * 64;
            const double r0 = rhs[lane * 2], r1 = rhs[lane * 2 + 1];
            double acc[8];
#pragma unroll
            for (int rr = 0; rr < 8; rr++) {
                const double2 m = *reinterpret_cast<const double2*>(blk + rr * 64 + lane * 2);
                acc[rr] = fma(m.x, r0, m.y * r1);
            }
#pragma unroll
            for (int rr = 0; rr < 8; rr++) {
                const double v = warp_sum(acc[rr]);
                if (lane == 0) w[kb * 64 + warp * 8 + rr] = v;
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < np; i += 256) {
        const double s = w[i];
        nv.s[o + i] = s;
        nv.a[o + i] = nv.bvec[o + i] - nv.Ws[o + i] * s;
    }
}

// Backward half of the solve (as k_trsv2<true>) by a CLUSTER of 4 CTAs per chain, for batches of about one chain per SM or
// less (the sampler's partial FULL calls, small batches): one CTA per chain then pays the latency of 12 dependent steps with a
// single CTA's loads in flight (141 chains: 0.71 ms per FULL estimate, the cluster: 0.45 ms; 18 chains: 141 vs 74 us per launch).
// A full batch is bandwidth bound and better off with one CTA per chain (154 against 172 us per launch): the host picks by the
// batch size (apm_ctx::trsv_cluster_max; launching both as a pair that decides on the device by the number of active chains
// costs more in empty cluster launches than the straggler rounds gain).  few_max: chains beyond it are not expected.  CTA r of
// the cluster reads rows 16 r .. 16 r + 15 of every 64-row block of the column block (2 of the 8 row groups), so that all
// of a step's loads of a thread are independent; the four 64-entry partial sums are exchanged through distributed shared
// memory (every CTA writes its vector into all four, one cluster barrier per step, buffers alternate by step parity) and
// every CTA finishes the step redundantly (same summation order in all four: identical s_k everywhere).
// grid 4 * chains, 256 threads, dynamic smem: w[np] + part[8][64] + rhs[64] + xch[2][4][64]
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(256) k_trsv_back_c4(const double* __restrict__ L, long long l_bs, int ld, int nb,
                                                                               const double* __restrict__ LinvT, long long inv_bs,
                                                                               NewtonVecs nv, int few_max) {
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) double smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int b = blockIdx.x >> 2;
    const int cr = (int)cluster.block_rank();
    if (b >= few_max || !nv.active[b] || nv.status[b] != 0) return;      // the same decision in the four CTAs of a cluster
    const int np = nb * 64;
    double* w = smem;             // [np]
    double* part = w + np;        // [8][64]
    double* rhs = part + 512;     // [64]
    double* xch = rhs + 64;       // [2][4][64]
    double* xch_remote[4];
#pragma unroll
    for (int q = 0; q < 4; q++) xch_remote[q] = cluster.map_shared_rank(xch, q);
    const double* Lb = L + (long long)b * l_bs;
    const double* Ib = LinvT + (long long)b * inv_bs;
    const long long o = (long long)b * nv.vs;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < np; i += 256) w[i] = nv.s[o + i];
    cluster.sync();               // all four CTAs are running (their shared memory may be written) and w is in place
    const int cgp = tid & 31, rg = tid >> 5;              // column pair, row slot 0..7
    const int rrow = 16 * cr + rg;                         // this thread's two rows inside every 64-row block: rrow, rrow + 8
    int par = 0;
    for (int kb = nb - 1; kb >= 0; kb--, par ^= 1) {
        {
            double a0 = 0.0, a1 = 0.0;
            const double* col = Lb + kb * 64 + cgp * 2;
            int ib = kb + 1;
            for (; ib + 3 < nb; ib += 4) {                 // 8 independent loads in flight
                double2 m[8];
#pragma unroll
                for (int q = 0; q < 8; q++) m[q] = *reinterpret_cast<const double2*>(col + (size_t)((ib + (q >> 1)) * 64 + rrow + 8 * (q & 1)) * ld);
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const double sr = w[(ib + (q >> 1)) * 64 + rrow + 8 * (q & 1)];
                    a0 = fma(m[q].x, sr, a0);
                    a1 = fma(m[q].y, sr, a1);
                }
            }
            for (; ib < nb; ib++) {
                const double2 m0 = *reinterpret_cast<const double2*>(col + (size_t)(ib * 64 + rrow) * ld);
                const double2 m1 = *reinterpret_cast<const double2*>(col + (size_t)(ib * 64 + rrow + 8) * ld);
                const double s0 = w[ib * 64 + rrow], s1 = w[ib * 64 + rrow + 8];
                a0 = fma(m0.x, s0, a0); a1 = fma(m0.y, s0, a1);
                a0 = fma(m1.x, s1, a0); a1 = fma(m1.y, s1, a1);
            }
            part[rg * 64 + cgp * 2] = a0;
            part[rg * 64 + cgp * 2 + 1] = a1;
        }
        __syncthreads();
        if (tid < 64) {
            double v = 0.0;
#pragma unroll
            for (int q = 0; q < 8; q++) v += part[q * 64 + tid];
#pragma unroll
            for (int q = 0; q < 4; q++) xch_remote[q][(par * 4 + cr) * 64 + tid] = v;
        }
        cluster.sync();
        if (tid < 64) {
            const double* x = xch + par * 256;
            rhs[tid] = w[kb * 64 + tid] - ((x[tid] + x[64 + tid]) + (x[128 + tid] + x[192 + tid]));
        }
        __syncthreads();
        {
            const double* blk = Ib + (size_t)kb * 4096 + (size_t)(warp * 8) * 64;
            const double r0 = rhs[lane * 2], r1 = rhs[lane * 2 + 1];
            double acc[8];
#pragma unroll
            for (int rr = 0; rr < 8; rr++) {
                const double2 m = *reinterpret_cast<const double2*>(blk + rr * 64 + lane * 2);
                acc[rr] = fma(m.x, r0, m.y * r1);
            }
#pragma unroll
            for (int rr = 0; rr < 8; rr++) {
                const double v = warp_sum(acc[rr]);
                if (lane == 0) w[kb * 64 + warp * 8 + rr] = v;
            }
        }
        __syncthreads();
    }
    // every CTA holds the whole solution: each writes a quarter
    for (int i = cr * 256 + tid; i < np; i += 1024) {
        const double sv = w[i];
        nv.s[o + i] = sv;
        nv.a[o + i] = nv.bvec[o + i] - nv.Ws[o + i] * sv;
    }
    cluster.sync();               // no CTA exits while another may still write its shared memory (none does after the last step; cheap)
}

// start of a Newton mode search: every chain active and (hybrid) in the B-space form; n_active = {B, 0, 0}
__global__ void k_newton_init(int* active, int* mask_b, int* mask_m, int* done_m, int* n_active, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) {
        active[b] = 1;
        if (mask_b) { mask_b[b] = 1; mask_m[b] = 0; done_m[b] = 0; }
    }
    if (b == 0) { n_active[0] = B; n_active[1] = 0; n_active[2] = 0; }
}

// diff = mean((fnew - f)^2); f <- fnew; iteration bookkeeping (lpa.py:96-102)
__global__ void __launch_bounds__(256) k_newton_finish(NewtonVecs nv) {
    __shared__ double red[8];
    const int b = blockIdx.x;
    if (!nv.active[b]) return;
    if (nv.status[b] != 0) {  // failed earlier (chol K / chol B): retire the chain
        if (threadIdx.x == 0) {
            nv.active[b] = 0;
            atomicSub(nv.n_active, 1);
            if (nv.done_m) nv.mask_b[b] = nv.mask_m[b] = 0;
        }
        return;
    }
    const long long o = (long long)b * nv.vs;
    double acc = 0.0;
    for (int i = threadIdx.x; i < nv.n; i += 256) {
        const double fn = nv.fnew[o + i];
        const double d = fn - nv.f[o + i];
        acc = fma(d, d, acc);
        nv.f[o + i] = fn;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; w++) tot += red[w];
        const double diff = tot / (double)nv.n;
        const int it = nv.iters[b] + 1;
        nv.iters[b] = it;
        bool still = true;
        if (!(diff == diff) || isinf(diff)) {
            nv.status[b] = 4;
            still = false;
        } else if (diff < nv.tol) {
            still = false;
        } else if (it >= nv.max_iters) {
            nv.status[b] = 2;
            still = false;
        }
        if (!still) {
            nv.active[b] = 0;
            atomicSub(nv.n_active, 1);
        }
        if (nv.done_m) {
            // Newton converges quadratically here (diff_{k+1} ~ 0.6-1.4 diff_k^2 on GP-probit data, profiles/): the next
            // iteration is predicted to be the chain's last one if pred_factor * diff^2 < tol.  A wrong guess only costs
            // time: an M-space iteration that does not finish is n^3/3 more work, a finish that was not predicted pays the
            // separate covariance phase -- the latter is the expensive one (latency-bound on few chains), hence factor < 1.
            const int was_m = nv.mask_m[b];
            if (!still) {
                nv.done_m[b] = was_m;
                nv.mask_b[b] = nv.mask_m[b] = 0;
                if (!was_m) atomicAdd(nv.n_active + 2, 1);
            } else {
                const int next_m = nv.pred_factor * diff * diff < nv.tol ? 1 : 0;
                nv.mask_m[b] = next_m;
                nv.mask_b[b] = 1 - next_m;
                if (next_m) atomicAdd(nv.n_active + 1, 1);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Expectation propagation for the probit likelihood (EXTENSION: the reference has no EP -- SURVEY App. D; the
// algorithm is Rasmussen & Williams, GPML, Alg. 3.5 with all sites updated from the same (mu, diag Sigma) before the
// posterior is recomputed -- "parallel EP" -- restated in the ep_approximation restatement under oracle/).
// Vector roles while EP runs (NewtonVecs): f = mu, W = tau~, Ws = sqrt(tau~), bvec = nu~, fnew = mu of the new sites.
// ------------------------------------------------------------------------------------------------
struct EpVecs {
    double* s2;        // diag(Sigma) [chain][np]
    double* delta;     // max |site change| of the last update, per chain
    double damping;
};

// s2 <- diag(K)
__global__ void k_ep_init(const double* __restrict__ K, long long k_bs, int ld, NewtonVecs nv, EpVecs ev) {
    const int b = blockIdx.x;
    const long long o = (long long)b * nv.vs;
    for (int i = threadIdx.x; i < nv.np; i += blockDim.x) ev.s2[o + i] = K[(long long)b * k_bs + (size_t)i * ld + i];
}

// cavity, probit moments, damped site update for every i at once (GPML eqs. 3.56, 3.58, 3.59)
__global__ void __launch_bounds__(256) k_ep_sites(NewtonVecs nv, EpVecs ev) {
    __shared__ double red[8];
    const int b = blockIdx.x;
    if (!nv.active[b] || nv.status[b] != 0) return;
    const long long o = (long long)b * nv.vs;
    double dmax = 0.0;
    for (int i = threadIdx.x; i < nv.np; i += 256) {
        double tau_n = 0.0, nu_n = 0.0;
        if (i < nv.n) {
            const double y = nv.y[i], mu = nv.f[o + i], s2 = ev.s2[o + i], tau = nv.W[o + i], nu = nv.bvec[o + i];
            const double tau_c = 1.0 / s2 - tau, nu_c = mu / s2 - nu;
            const double mu_c = nu_c / tau_c, s2_c = 1.0 / tau_c;
            const double den = sqrt(1.0 + s2_c);
            const double z = y * mu_c / den;
            const double r = exp(-0.5 * z * z - log_ndtr(z) - HALF_LOG_2PI);
            const double mu_h = mu_c + y * s2_c * r / den;
            const double s2_h = s2_c - s2_c * s2_c * r * (z + r) / (1.0 + s2_c);
            tau_n = tau + ev.damping * ((1.0 / s2_h - tau_c) - tau);
            nu_n = nu + ev.damping * ((mu_h / s2_h - nu_c) - nu);
            tau_n = fmax(tau_n, 0.0);
            const double d = fmax(fabs(tau_n - tau), fabs(nu_n - nu));
            dmax = (d == d) ? fmax(dmax, d) : CUDART_INF;   // NaN -> inf (reported as non-finite by k_ep_finish)
        }
        nv.W[o + i] = tau_n;
        nv.Ws[o + i] = sqrt(tau_n);
        nv.bvec[o + i] = nu_n;
    }
    dmax = warp_max(dmax);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dmax;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int w = 0; w < 8; w++) m = fmax(m, red[w]);
        ev.delta[b] = m;
    }
}

// s2_i = K_ii - sum_j Z_ij^2 with Z = K S^1/2 L_B^{-T} (diag of Sigma = K - Z Z^T); warp per row; grid (np/32, chains)
__global__ void __launch_bounds__(256) k_ep_diag_sigma(const double* __restrict__ K, long long k_bs, const double* __restrict__ Z,
                                                       long long z_bs, int ld, NewtonVecs nv, EpVecs ev) {
    const int b = blockIdx.y;
    if (!nv.active[b] || nv.status[b] != 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const int r = blockIdx.x * 32 + warp * 4 + rr;
        const double* row = Z + (long long)b * z_bs + (size_t)r * ld;
        double acc = 0.0;
        for (int j = lane * 2; j < nv.np; j += 64) {
            const double2 m = *reinterpret_cast<const double2*>(row + j);
            acc = fma(m.x, m.x, acc);
            acc = fma(m.y, m.y, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) ev.s2[(long long)b * nv.vs + r] = K[(long long)b * k_bs + (size_t)r * ld + r] - acc;
    }
}

// mu <- mu_new; iteration bookkeeping; a chain is done when the largest site change fell below tol
__global__ void k_ep_finish(NewtonVecs nv, EpVecs ev) {
    const int b = blockIdx.x;
    if (!nv.active[b]) return;
    const long long o = (long long)b * nv.vs;
    if (nv.status[b] == 0)
        for (int i = threadIdx.x; i < nv.np; i += blockDim.x) nv.f[o + i] = nv.fnew[o + i];
    if (threadIdx.x == 0) {
        bool still = true;
        if (nv.status[b] != 0) {
            still = false;
        } else {
            const double d = ev.delta[b];
            const int it = nv.iters[b] + 1;
            nv.iters[b] = it;
            if (!(d == d) || isinf(d)) {
                nv.status[b] = 4;
                still = false;
            } else if (d < nv.tol) {
                still = false;
            } else if (it >= nv.max_iters) {
                nv.status[b] = 2;
                still = false;
            }
        }
        if (!still) {
            nv.active[b] = 0;
            atomicSub(nv.n_active, 1);
        }
    }
}

// approximate log marginal likelihood (lpa.py:105-106): -a.f/2 + sum log Phi(y f) - sum log diag L
__global__ void __launch_bounds__(256) k_laplace_lml(NewtonVecs nv, const double* logdet_parts, int ld_stride, int nb,
                                                     double* lml_out) {
    __shared__ double red[2][8];
    const int b = blockIdx.x;
    if (nv.status[b] != 0) {
        if (threadIdx.x == 0) lml_out[b] = nan("");
        return;
    }
    const long long o = (long long)b * nv.vs;
    double af = 0.0, ll = 0.0;
    for (int i = threadIdx.x; i < nv.n; i += 256) {
        const double f = nv.f[o + i];
        af = fma(nv.a[o + i], f, af);
        ll += log_ndtr(nv.y[i] * f);
    }
    af = warp_sum(af);
    ll = warp_sum(ll);
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = af;
        red[1][threadIdx.x >> 5] = ll;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0.0, s1 = 0.0, ld = 0.0;
        for (int w = 0; w < 8; w++) {
            s0 += red[0][w];
            s1 += red[1][w];
        }
        for (int k = 0; k < nb; k++) ld += logdet_parts[(size_t)b * ld_stride + k];
        lml_out[b] = -0.5 * s0 + s1 - ld;
    }
}

// ------------------------------------------------------------------------------------------------
// Importance-sampling epilogue (estimators.py:225-240) in two stages (k_is_logw: a warp per sample; k_is_epilogue: the
// log-sum-exp of a chain):
//   lw_s = sum_i log Phi(y_i F_si) - q_K/2 - sum log diag L_K + q_u/2 + sum log diag L_C
//   with q_K = |L_K^{-1} f_s|^2 (rows of Zf) and q_u = |u_s|^2 (== (f_s-mu)^T C^{-1} (f_s-mu), est.py:232-234)
//   out = logsumexp_s lw_s - log N
// mode 1 (prior MC, estimators.py:323-325): lw_s = sum_i log Phi(y_i F_si) only.
// ------------------------------------------------------------------------------------------------
struct EpilogueParams {
    const double* F; const double* Zf; const double* UT; long long bs; int ld;   // [chain][Npad][np]
    const double* y; int n, N;
    const double* logdetK; const double* logdetC; int ld_stride; int nb; const int* slot_idx;   // per slot parts
    const int* status;
    double* logml; double* logw;   // logw: [chain][N] log-weights (always written; workspace of the two stages)
    int mode;
    const double* mt; long long mt_bs;   // optional [slot][np]: L_K^{-1} f_s = mt + Zf row (factored cache, see run_is_tail)
};

// stage 1: one warp per (chain, sample): lw_s -> p.logw[chain][s].  grid (ceil(N / 8), chains), 256 threads.
// The element loop is unrolled by 4 with all loads issued first (the kernel is memory-latency bound otherwise); every
// lane still accumulates its elements in increasing order.
__global__ void __launch_bounds__(256) k_is_logw(EpilogueParams p) {
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.x * 8 + warp;
    if (s >= p.N) return;
    if (p.status && p.status[b] != 0) {      // failed chain: its log-weights are NaN, never stale workspace contents
        if (lane == 0) p.logw[(size_t)b * p.N + s] = nan("");
        return;
    }
    const double* Fr = p.F + (long long)b * p.bs + (size_t)s * p.ld;
    double ll = 0.0, qk = 0.0, qu = 0.0;
    if (p.mode == 0) {
        const double* Zr = p.Zf + (long long)b * p.bs + (size_t)s * p.ld;
        const double* Ur = p.UT + (long long)b * p.bs + (size_t)s * p.ld;
        const double* mt = p.mt ? p.mt + chain_index(p.slot_idx, b) * p.mt_bs : nullptr;
        int i = lane;
        for (; i + 96 < p.n; i += 128) {
            double f[4], z[4], u[4], yy[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                f[q] = Fr[i + 32 * q];
                z[q] = Zr[i + 32 * q];
                u[q] = Ur[i + 32 * q];
                yy[q] = p.y[i + 32 * q];
                if (mt) z[q] += mt[i + 32 * q];
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                ll += log_ndtr(yy[q] * f[q]);
                qk = fma(z[q], z[q], qk);
                qu = fma(u[q], u[q], qu);
            }
        }
        for (; i < p.n; i += 32) {
            ll += log_ndtr(p.y[i] * Fr[i]);
            const double z = mt ? mt[i] + Zr[i] : Zr[i], u = Ur[i];
            qk = fma(z, z, qk);
            qu = fma(u, u, qu);
        }
    } else {
        int i = lane;
        for (; i + 96 < p.n; i += 128) {
            double f[4], yy[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                f[q] = Fr[i + 32 * q];
                yy[q] = p.y[i + 32 * q];
            }
#pragma unroll
            for (int q = 0; q < 4; q++) ll += log_ndtr(yy[q] * f[q]);
        }
        for (; i < p.n; i += 32) ll += log_ndtr(p.y[i] * Fr[i]);
    }
    ll = warp_sum(ll);
    qk = warp_sum(qk);
    qu = warp_sum(qu);
    if (lane == 0) {
        double v = ll;
        if (p.mode == 0) {
            double ldK = 0.0, ldC = 0.0;
            const long long sl = chain_index(p.slot_idx, b);
            for (int k = 0; k < p.nb; k++) {
                ldK += p.logdetK[(size_t)sl * p.ld_stride + k];
                ldC += p.logdetC[(size_t)sl * p.ld_stride + k];
            }
            v = ll + (-0.5 * qk - ldK) - (-0.5 * qu - ldC);
        }
        p.logw[(size_t)b * p.N + s] = v;
    }
}

// stage 2: logml[chain] = logsumexp_s lw_s - log N (estimators.py:240), one CTA per chain
__global__ void __launch_bounds__(256) k_is_epilogue(EpilogueParams p) {
    __shared__ double red[8];
    const int b = blockIdx.x;
    if (p.status && p.status[b] != 0) {
        if (threadIdx.x == 0) p.logml[b] = nan("");
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double* lw = p.logw + (size_t)b * p.N;
    double m = -INFINITY;
    for (int s = threadIdx.x; s < p.N; s += 256) m = fmax(m, lw[s]);
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
    for (int w = 1; w < 8; w++) m = fmax(m, red[w]);
    __syncthreads();
    if (m == -INFINITY) {                    // every weight is zero: scipy's logsumexp gives -inf, not NaN (estimators.py:240)
        if (threadIdx.x == 0) p.logml[b] = -INFINITY;
        return;
    }
    double sum = 0.0;
    for (int s = threadIdx.x; s < p.N; s += 256) sum += exp(lw[s] - m);
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; w++) tot += red[w];
        p.logml[b] = (log(tot) + m) - log((double)p.N);
    }
}

// ------------------------------------------------------------------------------------------------
// layout helpers
// ------------------------------------------------------------------------------------------------
// Anti-transpose of a lower-triangular matrix: dst[i][j] = src[np-1-j][np-1-i] (i >= j), zeros above the diagonal inside
// the 64x64 diagonal blocks.  Maps L' = chol(P M P) to V = U^T with M = U U^T (V lower triangular, V^T V... see
// run_covariance_factored) and back (it is an involution).  32x32 tiles through shared memory; grid (np/32, np/32, chains).
__global__ void k_antitranspose(const double* __restrict__ src, long long s_bs, const int* s_idx, double* __restrict__ dst,
                                long long d_bs, const int* d_idx, int np, const int* status) {
    __shared__ double tile[32][33];
    const int b = blockIdx.z;
    if (status && status[b] != 0) return;
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    if (i0 < j0 && (i0 >> 6) != (j0 >> 6)) return;            // strictly upper, outside a diagonal block: never read
    const double* S = src + chain_index(s_idx, b) * s_bs;
    double* Dm = dst + chain_index(d_idx, b) * d_bs;
    const int tx = threadIdx.x, ty = threadIdx.y;             // 32 x 8
    const bool upper = i0 < j0;
    if (!upper) {
        const int r0 = np - 32 - j0, c0 = np - 32 - i0;
        for (int r = ty; r < 32; r += 8) tile[r][tx] = S[(size_t)(r0 + r) * np + c0 + tx];
    }
    __syncthreads();
    for (int a = ty; a < 32; a += 8) {
        const int i = i0 + a, j = j0 + tx;
        Dm[(size_t)i * np + j] = (!upper && i >= j) ? tile[31 - tx][31 - a] : 0.0;
    }
}

// out[slot][j] = sum_{i >= j} L[i][j] x[i]   (L^T x for lower-triangular L); grid (nb, chains), 256 threads = 4 row groups
// reverse_out: the result is written index-reversed (out[np-1-j]), the right-hand side of the reversed system M' x' = g'
__global__ void __launch_bounds__(256) k_lt_matvec(const double* __restrict__ L, long long l_bs, const int* l_idx, int ld, int nb,
                                                   const double* __restrict__ x, long long x_bs, double* __restrict__ out,
                                                   long long o_bs, const int* o_idx, const int* status, const int* mask,
                                                   int reverse_out) {
    __shared__ double part[4][64];
    const int b = blockIdx.y, jb = blockIdx.x;
    if ((status && status[b] != 0) || (mask && !mask[b])) return;
    const double* Lb = L + chain_index(l_idx, b) * l_bs;
    const double* xb = x + (long long)b * x_bs;
    const int c = threadIdx.x & 63, rg = threadIdx.x >> 6;
    double a0 = 0.0, a1 = 0.0;
    const int np = nb * 64;
    int i = jb * 64 + rg;
    for (; i + 4 < np; i += 8) {
        a0 = fma(Lb[(size_t)i * ld + jb * 64 + c], xb[i], a0);
        a1 = fma(Lb[(size_t)(i + 4) * ld + jb * 64 + c], xb[i + 4], a1);
    }
    for (; i < np; i += 4) a0 = fma(Lb[(size_t)i * ld + jb * 64 + c], xb[i], a0);
    part[rg][c] = a0 + a1;
    __syncthreads();
    if (threadIdx.x < 64) {
        const int j = jb * 64 + c;
        out[chain_index(o_idx, b) * o_bs + (reverse_out ? np - 1 - j : j)] = (part[0][c] + part[1][c]) + (part[2][c] + part[3][c]);
    }
}

// M-space Newton step, last part: mu~[j] = s'[np-1-j] (solution of the reversed system), f_new = L_K mu~ (lower
// triangular mat-vec, warp per row), mu~ stored in the slot (it is L_K^-1 f_new, what the factored cache keeps).
// grid (np/32, chains), 256 threads = 8 warps x 4 rows
__global__ void __launch_bounds__(256) k_l_matvec_rev(const double* __restrict__ L, long long l_bs, const int* l_idx, int ld, int np,
                                                      const double* __restrict__ srev, long long s_bs, double* __restrict__ fnew,
                                                      long long f_bs, double* __restrict__ mt, long long mt_bs, const int* mt_idx,
                                                      const int* status, const int* mask) {
    const int b = blockIdx.y;
    if ((status && status[b] != 0) || (mask && !mask[b])) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double* Lb = L + chain_index(l_idx, b) * l_bs;
    const double* sb = srev + (long long)b * s_bs;
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const int r = blockIdx.x * 32 + warp * 4 + rr;
        const double* row = Lb + (size_t)r * ld;
        double acc = 0.0;
        for (int j = lane; j <= r; j += 32) acc = fma(row[j], sb[np - 1 - j], acc);
        acc = warp_sum(acc);
        if (lane == 0) {
            fnew[(long long)b * f_bs + r] = acc;
            mt[chain_index(mt_idx, b) * mt_bs + r] = sb[np - 1 - r];
        }
    }
}

// dst[b] = !src[b]
__global__ void k_mask_not(const int* src, int* dst, int n) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n) dst[b] = src[b] ? 0 : 1;
}

// pad region of a [np][np] matrix -> identity (after importing an n x n matrix)
__global__ void k_pad_identity(double* M, long long bs, int n, int np) {
    double* Mb = M + (long long)blockIdx.y * bs;
    const int npad = np - n;
    const long long total = (long long)np * npad * 2;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int r, c;
        if (e < (long long)np * npad) {  // right strip: all rows, cols n..np
            r = (int)(e / npad);
            c = n + (int)(e % npad);
        } else {  // bottom strip: rows n..np, all cols
            const long long e2 = e - (long long)np * npad;
            r = n + (int)(e2 / np);
            c = (int)(e2 % np);
        }
        Mb[(size_t)r * np + c] = (r == c) ? 1.0 : 0.0;
    }
}
// out[r][c] (n x n, dense) = r >= c ? M[r][c] : (mirror ? M[c][r] : 0)
__global__ void k_export_lower(const double* __restrict__ M, int np, int n, double* __restrict__ out, int mirror) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c >= n) return;
    double v;
    if (r >= c) v = M[(size_t)r * np + c];
    else v = mirror ? M[(size_t)c * np + r] : 0.0;
    out[(size_t)r * n + c] = v;
}
__global__ void k_fill_int(int* p, int v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_fill_double(double* p, double v, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// copy a [np] vector per chain with optional slot indirection on either side
__global__ void k_copy_vec(const double* src, long long s_bs, const int* s_idx, double* dst, long long d_bs,
                           const int* d_idx, int len, const int* status) {
    const int b = blockIdx.y;
    if (status && status[b] != 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) dst[chain_index(d_idx, b) * d_bs + i] = src[chain_index(s_idx, b) * s_bs + i];
}

// fp64 peak probes (bench.py roofline denominators): dependent-free DMMA / DFMA streams
__global__ void k_peak_dmma(double* out, int iters) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-9, bb = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) dmma884(c[i][0], c[i][1], a, bb);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}
// DMMA throughput as a function of resident warps: NACC independent accumulators per warp
template <int NACC>
__global__ void k_peak_dmma_n(double* out, int iters) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-9, bb = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma884(c[i][0], c[i][1], a, bb);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}
__global__ void k_peak_dfma(double* out, int iters) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = threadIdx.x * 1e-3 + i;
    const double a = 1.0000001, bb = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) c[i] = fma(c[i], a, bb);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
    if (s == 123.456) out[0] = s;
}

}  // namespace apm
