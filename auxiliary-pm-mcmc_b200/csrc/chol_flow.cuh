// chol_flow.cuh -- batched blocked fp64 Cholesky as ONE persistent, warp-specialised dataflow kernel (sm_100a).
//
// Replaces LAPACK dpotrf of the reference's hot path (lpa.py:92; estimators.py:206, 209) for a batch of chains.
//
// Structure (per CTA: 1 producer warp + 4 consumer warps, 3 CTAs per SM):
//   * tasks of the left-looking blocked factorisation on 64x64 tiles, claimed from a global queue in dependency order:
//       diag(k)     : L_kk = chol(A_kk - sum_{j<k} L_kj L_kj^T)
//       panel(k, i) : L_ik = (A_ik - sum_{j<k} L_ij L_kj^T) L_kk^{-T},  i > k
//     A = diag(scale) * src * diag(scale) (+ I): B = I + W^1/2 K W^1/2 (lpa.py:91) is never stored.
//   * the PRODUCER lane claims the next task while the consumers still work on the current one, streams the source tile
//     (no dependency) and then, after polling the per-row progress counters (acquire loads), the GEMM operands as 64 x 16
//     fp64 boxes (8 KB, one 128-byte swizzle row per matrix row) with TMA (cp.async.bulk.tensor.2d + mbarrier complete_tx)
//     into a ring of stages; the
//     packed diagonal block L_kk (36 lower 8x8 blocks + the 8 inverses of its diagonal 8x8 blocks, 22.5 KB contiguous in
//     global memory, written by diag(k)) arrives by one 1-D bulk copy.  Dependencies are checked per operand block, so a
//     panel GEMM runs before L_kk exists and a diagonal GEMM runs ahead of the last panel of its row (look-ahead for free).
//   * the CONSUMER warps own 16 rows x 64 columns of the tile as DMMA accumulators (2 x 8 m8n8 tiles) from the source-tile
//     load to the final store: GEMM from the swizzled stages (full/empty mbarriers, no block barrier in the k-loop), then
//     the triangular solve entirely in registers on the tensor pipe.  The m8n8k4 contraction index is a free
//     permutation as long as A and B agree: with k -> columns {2t, 2t+1} the A fragment of X * M IS the lane's own pair of
//     accumulator values, so   X_p = T_p * inv(L_pp)^T   and   T_q -= X_p * L_qp^T   need no shuffle and no shared-memory
//     round trip; B fragments are single 16-byte loads from the packed 8x8 blocks.
//   * diag(k): 8 column panels; the warp owning the 8x8 diagonal block factors AND inverts it with quad shuffles
//     (lane (g,t) holds D[g][2t..2t+1]), publishes both through the packed block buffer, every warp solves its rows on the
//     tensor pipe and publishes them as the B operand of the trailing update: 2 named barriers per panel.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace apm {

constexpr int CF_CONSUMER_WARPS = 4;
constexpr int CF_CONSUMERS = CF_CONSUMER_WARPS * 32;
constexpr int CF_THREADS = CF_CONSUMERS + 32;
constexpr int CF_KC = 16;                              // doubles per k-chunk = one 128-byte swizzle row
constexpr int CF_CHUNK_BYTES = TB * CF_KC * 8;         // 8192: one operand box
#ifndef APM_CF_STAGES
#define APM_CF_STAGES 3
#endif
#ifndef APM_CF_MIN_CTAS
#define APM_CF_MIN_CTAS 3
#endif
constexpr int CF_STAGES = APM_CF_STAGES;
constexpr int DP_LBLOCKS = 36;                         // lower 8x8 blocks (q >= p) of a 64x64 triangle, block (q,p) at q(q+1)/2 + p
constexpr int DP_DOUBLES = (DP_LBLOCKS + 8) * 64;      // + inverses of the 8 diagonal 8x8 blocks
constexpr int DP_BYTES = DP_DOUBLES * 8;               // 22528
// control block: barriers / task slots (512 B), W_r or y and b_r slices of the stages (128 B per stage each), two right-hand-side blocks
constexpr int CF_WST_OFF = 512, CF_BST_OFF = CF_WST_OFF + CF_STAGES * 128, CF_RHS_OFF = CF_BST_OFF + CF_STAGES * 128;
constexpr int CF_CTRL_BYTES = CF_RHS_OFF + 2 * 64 * 8;
constexpr int CF_SMEM_BYTES = 1024 + CF_STAGES * 2 * CF_CHUNK_BYTES + DP_BYTES + CF_CTRL_BYTES;   // 1024: manual alignment slack

__device__ __forceinline__ int dp_block(int q, int p) { return (q * (q + 1) / 2 + p) * 64; }
__device__ __forceinline__ int dp_inv(int p) { return (DP_LBLOCKS + p) * 64; }

// ---- mbarrier / TMA / proxy-fence wrappers (PTX ISA 8.x, sm_90+) ---------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "APM_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra APM_MBAR_DONE;\n"
        "bra APM_MBAR_WAIT;\n"
        "APM_MBAR_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
                 "l"(tm), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
// generic <-> async proxy ordering.  The unqualified form only fences the shared-memory view (SASS FENCE.VIEW.ASYNC.S): data
// that crosses CTAs through GLOBAL memory and is read or written by TMA / bulk copies needs the .global form
// (FENCE.VIEW.ASYNC.G) on both sides of the release / acquire of its flag.
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;\n" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }

__device__ __forceinline__ int cf_ld_relaxed(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void cf_st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void cf_consumer_bar() { asm volatile("bar.sync 1, %0;\n" ::"n"(CF_CONSUMERS) : "memory"); }
// publish barrier of a panel task: only warp 0 (which releases the progress counter) waits, the others just arrive.  Two ids,
// alternating by task parity: a warp can be at most one task ahead of warp 0.
__device__ __forceinline__ void cf_publish_arrive(int odd) {
    if (odd) asm volatile("bar.arrive 3, %0;\n" ::"n"(CF_CONSUMERS) : "memory");
    else asm volatile("bar.arrive 2, %0;\n" ::"n"(CF_CONSUMERS) : "memory");
}
__device__ __forceinline__ void cf_publish_sync(int odd) {
    if (odd) asm volatile("bar.sync 3, %0;\n" ::"n"(CF_CONSUMERS) : "memory");
    else asm volatile("bar.sync 2, %0;\n" ::"n"(CF_CONSUMERS) : "memory");
}

// ---- parameters ----------------------------------------------------------------------------------------------------
struct CholFlowParams {
    const double* src; long long src_bs; int lds; const int* src_idx;
    double* dst; long long dst_bs; int ldd; const int* dst_idx;
    int dst_m0;                  // tensor-map matrix index of chain 0 when dst_idx == null (lane views: offset into the root buffer)
    int src_m0;                  // same for the source tensor map when src_idx == null
    int np;                      // rows per matrix in the tensor maps (n padded)
    int zero;                    // always 0, but only known at run time (see cf_release_stage)
    // k_chol_flow<true>: the source matrix is M' = P (I + L_K^T W L_K) P and is never stored: its tiles are accumulated on the
    // fly from k-major 16x16 boxes of L_K (tensor map tml) scaled by W before the factorisation's own k-loop (src unused)
    const int* lk_idx; const double* w; long long w_bs;
    // k_chol_flow<true> only: optional second copy of the factor, anti-transposed (vt[np-1-c][np-1-r] = L'[r][c]) into matrix
    // lk_idx[b] of vt_out -- V = U^T of M = U U^T, the operand the importance-sampling tail solves with (no transpose kernel)
    double* vt_out; long long vt_bs;
    // Optional fused forward substitution y = L^-1 t (the first half of the Newton step's cho_solve, lpa.py:94): diag(k) adds
    // up L_k,0..k-1 y_0..k-1 from the operand chunks it streams anyway and solves its 64 rows with the packed L_kk once the
    // factor is published.  fwd_t: right-hand side, fwd_y: result (both [chain][np], stride fwd_bs), yprog[chain]: finished blocks.
    const double* fwd_t; double* fwd_y; long long fwd_bs; int* yprog;
    // k_chol_flow<true, true> only: fwd_t is not read -- the right-hand side t' = P L_K^T b is accumulated by the diagonal tasks
    // from L_K (which they stream for M' anyway) and b = lt_b [chain][np] (stride fwd_bs)
    const double* lt_b;
    const double* scale; long long scale_bs;
    int add_identity;
    int nb;
    double* logdet_parts; int logdet_stride; const int* logdet_idx;
    double* inv_out; long long inv_bs;          // optional (L_kk^{-1})^T of every diagonal block, [chain][nb][64*64]
    int* status; int fail_code;
    const int* active;
    int nchains;
    int* counter;                // [0] task queue head, [1] number of chains to factorise (both written by k_chol_flow_init)
    int* progress;               // [nchains][nb] finished column blocks per block row
    int* list;                   // [nchains] compacted indices of the chains to factorise (status == 0 and active)
    double* diagpack;            // [nchains][nb][DP_DOUBLES]
    int spin_ns;
};

// Zero the queue head and the progress counters, and compact the chains to factorise (status == 0 and inside the optional
// Newton mask) into an ordered list: a launch for a few straggler chains enumerates only their tasks.  One launch instead
// of two memsets + a snapshot kernel; block 0 / warp 0 does the (ballot) compaction.
__global__ void k_chol_flow_init(int* counter, int* progress, int* list, const int* status, const int* active, int nchains, int nb,
                                 unsigned long long* work, unsigned long long* work2, int* yprog = nullptr) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < nchains * nb) progress[e] = 0;
    if (yprog && e < nchains) yprog[e] = 0;
    if (blockIdx.x != 0) return;
    // ordered compaction by block 0 (256 threads = 8 warps): every pass handles 256 chains with one round of independent
    // loads (a single warp walking the chains paid one dependent global load per 32 chains)
    __shared__ int wcount[8];
    __shared__ int base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nchains; b0 += 256) {
        const int b = b0 + threadIdx.x;
        const bool on = b < nchains && status[b] == 0 && (!active || active[b]);
        const unsigned m = __ballot_sync(0xffffffffu, on);
        if (lane == 0) wcount[warp] = __popc(m);
        __syncthreads();
        int off = base;
        for (int w = 0; w < warp; w++) off += wcount[w];
        if (on) list[off + __popc(m & ((1u << lane) - 1u))] = b;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 8; w++) t += wcount[w];
            base += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int count = base;
        counter[0] = 0;
        counter[1] = count;
        if (work && count) atomicAdd(work, (unsigned long long)count);   // chain-Choleskys executed (work accounting)
        if (work2 && count) atomicAdd(work2, (unsigned long long)count); // ... and M' builds, when the source is the fused SYRK
    }
}

struct CfTask { int type, k, b, i; };   // type 0: panel(k, i), 1: diag(k), -1: no more tasks

// ---- consumer-side pieces -----------------------------------------------------------------------------------------
// accumulators of one consumer warp: rows 16*warp + 8*mt + g, columns 8*nt + 2t (+1)
typedef double CfAcc[2][8][2];

// Source tile from the ring: the producer streams the 64x64 tile as four 64x16 boxes (two stages, A-half / B-half each) ahead
// of the GEMM chunks -- it never depends on another task, so it is in shared memory before the consumers get to the task.
// Box j holds columns 16j..16j+15; this thread's pair (row r, columns 8nt+2t, +1) is the 16-byte segment 4(nt&1)+t of row r
// in box nt/2, stored at segment position (4(nt&1)+t) ^ (r & 7).   acc <- diag(rs) * S * diag(cs) (+ I).
__device__ __forceinline__ void cf_src_from_stage(CfAcc& acc, const unsigned char* stage, int half, const double* rs, const double* cs,
                                                  bool add_identity, int warp, int g, int t) {
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
        const int r = warp * 16 + mt * 8 + g;
        const double rsv = rs ? rs[r] : 1.0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int nt = half * 4 + q;                     // this stage covers column tiles 4*half .. 4*half+3
            const int c = nt * 8 + 2 * t;
            const unsigned char* box = stage + (q >> 1) * CF_CHUNK_BYTES;
            double2 v = *reinterpret_cast<const double2*>(box + r * 128 + ((((q & 1) * 4 + t) ^ g) << 4));
            if (rs || cs) {
                v.x = rsv * v.x * (cs ? cs[c] : 1.0);
                v.y = rsv * v.y * (cs ? cs[c + 1] : 1.0);
            }
            if (add_identity) {
                if (r == c) v.x += 1.0;
                if (r == c + 1) v.y += 1.0;
            }
            if (half == 0) { acc[mt][q][0] = v.x; acc[mt][q][1] = v.y; }
            else { acc[mt][4 + q][0] = v.x; acc[mt][4 + q][1] = v.y; }
        }
    }
}

// Hand a ring stage back to the producer.  The stage may only be overwritten (by TMA, asynchronously) after every fragment
// load of this warp has completed; a load has certainly completed once an instruction that consumes its result has issued.
// mbarrier.arrive has no data dependency on the loads, so the hardware may perform it while they are still queued (measured:
// wrong tiles as soon as two CTAs share an SM).  The arrive therefore gets an address operand that depends on `dep`, a value
// computed from the results of the chunk's last DMMAs / the registers the loads filled: (dep & zero) == 0 at run time.
__device__ __forceinline__ void cf_release_stage(uint32_t bar, uint32_t dep, int zero, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar + (dep & (uint32_t)zero));
}
__device__ __forceinline__ uint32_t cf_acc_dep(const CfAcc& acc) {
    uint32_t d = 0;
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < 8; nt++) d |= (uint32_t)__double2hiint(acc[mt][nt][1]);
    return d;
}

// acc -= A(rows of this warp) * B(all 64 rows)^T over one 16-deep chunk.  Stage layout: row r at r*128 bytes, its 16-byte
// segment c at position c ^ (r & 7) (TMA SWIZZLE_128B).  DMMA kk contracts the columns {2kk, 2kk+1, 2kk+8, 2kk+9}: lane t
// reads segment kk ^ 4(t>>1), half t&1 -- the 16 lanes of a half-warp then hit 16 distinct 8-byte bank pairs.
// FWD: ys[mt] -= (row of A) . y over this lane's four columns of the chunk (ych: the 16 entries of y that go with the chunk);
// the A fragment of DMMA kk is column 2 kk + (t & 1) + 8 (t >> 1) of the chunk.
template <bool DIAG, bool FWD = false>
__device__ __forceinline__ void cf_gemm_chunk(CfAcc& acc, const unsigned char* sA, const unsigned char* sB, int warp, int g, int t,
                                              const double* ych = nullptr, double* ys = nullptr) {
    const uint32_t lo0 = (uint32_t)(((((t >> 1) << 2) ^ g) << 4) | ((t & 1) << 3));
    const unsigned char* a_base = sA + (warp * 16 + g) * 128;
    const unsigned char* b_base = sB + g * 128;
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
        const uint32_t lo = lo0 ^ (uint32_t)(kk << 4);
        double a[2], b[8];
#pragma unroll
        for (int mt = 0; mt < 2; mt++) a[mt] = -*reinterpret_cast<const double*>(a_base + mt * 1024 + lo);
        if (FWD) {
            const double yv = ych[2 * kk + (t & 1) + 8 * (t >> 1)];
            ys[0] = fma(a[0], yv, ys[0]);
            ys[1] = fma(a[1], yv, ys[1]);
        }
#pragma unroll
        for (int nt = 0; nt < 8; nt++)
            if (!DIAG || nt <= 2 * warp + 1) b[nt] = *reinterpret_cast<const double*>(b_base + nt * 1024 + lo);
#pragma unroll
        for (int nt = 0; nt < 8; nt++)
            if (!DIAG || nt <= 2 * warp + 1) {
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
                    if (!DIAG || nt <= 2 * warp + mt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
            }
    }
}

// Source tiles of M' = P (I + L_K^T W L_K) P on the fly ("TN" chunk: sum over ROWS r of the row-major L_K).  Tile (i, k) of M'
// is tile (I, K) = (nb-1-i, nb-1-k) of M with both index orders reversed:
//   S[m'][n'] = [i == k][m' == n'] + sum_{r >= 64 K} W_r L_K[r][64 I + 63 - m'] L_K[r][64 K + 63 - n'].
// A stage holds 16 rows r as eight 16x16 boxes (k-major: row r at r*128 bytes inside a box, 16-byte segment c at c ^ (r & 7)):
// boxes 0..3 = column block I (A operand), 4..7 = column block K (B operand), plus W_r of the 16 rows.  This warp's output
// rows m' = 16 warp + 8 mt + g read source column 63 - m' = box 3 - warp, in-box column 15 - 8 mt - g; DMMA kk contracts the
// rows {2t + (kk & 1) + 8 (kk >> 1)}: their (r & 7) in {0,2,4,6} (+1) spreads the 16 lanes of a half-warp over 16 bank pairs.
// LT (diagonal tasks of a Newton M-space round): the A fragment of a diagonal task is L_K[r][64 K + 63 - n'] for this lane's
// output index n' = 16 warp + 8 mt + g -- exactly the terms of t'[64 k + n'] = sum_r L_K[r][.] b_r (the reversed L_K^T b, right-hand
// side of the round's solve): tq[mt] collects them with b_r from sBv (the 16 rows' entries of b), k_lt_matvec is not launched.
template <bool DIAG, bool LT = false>
__device__ __forceinline__ void cf_gemm_chunk_tn(CfAcc& acc, const unsigned char* sA, const unsigned char* sB, const double* sW,
                                                 int warp, int g, int t, const double* sBv = nullptr, double* tq = nullptr) {
    const unsigned char* a_box = sA + (3 - warp) * 2048;
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
        const int kr = 8 * (kk >> 1) + 2 * t + (kk & 1);
        const double wk = sW[kr];
        const uint32_t row = (uint32_t)kr * 128u, x = (uint32_t)(kr & 7);
        double a[2], b[8];
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            const uint32_t mm = (uint32_t)(15 - 8 * mt - g);
            const double raw = *reinterpret_cast<const double*>(a_box + row + (((mm >> 1) ^ x) << 4) + ((mm & 1) << 3));
            a[mt] = wk * raw;
            if (LT) tq[mt] = fma(raw, sBv[kr], tq[mt]);
        }
#pragma unroll
        for (int nt = 0; nt < 8; nt++)
            if (!DIAG || nt <= 2 * warp + 1) {
                const uint32_t mm = (uint32_t)(15 - 8 * (nt & 1) - g);
                b[nt] = *reinterpret_cast<const double*>(sB + (3 - (nt >> 1)) * 2048 + row + (((mm >> 1) ^ x) << 4) + ((mm & 1) << 3));
            }
#pragma unroll
        for (int nt = 0; nt < 8; nt++)
            if (!DIAG || nt <= 2 * warp + 1) {
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
                    if (!DIAG || nt <= 2 * warp + mt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
            }
    }
}

// acc <- acc * L^{-T} with L packed in dp (blocks + diagonal-block inverses); all in registers on the tensor pipe.
// SKIP_ZERO: acc is the identity (rows above a panel are structurally zero there): panels p < row tile are skipped.
template <bool SKIP_ZERO>
__device__ __forceinline__ void cf_trsm_regs(CfAcc& acc, const double* dp, int warp, int g, int t) {
    const int off = g * 8 + 2 * t;
#pragma unroll
    for (int p = 0; p < 8; p++) {
        if (SKIP_ZERO && p < 2 * warp) continue;
        const double2 bi = *reinterpret_cast<const double2*>(dp + dp_inv(p) + off);
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            double x0 = 0.0, x1 = 0.0;
            dmma884(x0, x1, acc[mt][p][0], bi.x);
            dmma884(x0, x1, acc[mt][p][1], bi.y);
            acc[mt][p][0] = x0;
            acc[mt][p][1] = x1;
        }
#pragma unroll
        for (int q = p + 1; q < 8; q++) {
            const double2 bl = *reinterpret_cast<const double2*>(dp + dp_block(q, p) + off);
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
                dmma884(acc[mt][q][0], acc[mt][q][1], -acc[mt][p][0], bl.x);
                dmma884(acc[mt][q][0], acc[mt][q][1], -acc[mt][p][1], bl.y);
            }
        }
    }
}

// Cholesky factor AND inverse of an 8x8 SPD block distributed over a warp in DMMA accumulator layout: lane (g, t) holds
// D[g][2t], D[g][2t+1] in (d0, d1).  Returns L (strict upper zeroed) in (d0, d1) and L^-1 in (y0, y1).  Right-looking, one
// column per step; communication by shuffles only.  L_jj = piv * rsqrt(piv) as in the round-1 engine.  One copy of the code
// (noinline): it runs 8 times per diagonal task on one warp.
struct CfChol8 { double d0, d1, y0, y1; int bad; };
__device__ __noinline__ CfChol8 cf_chol8_inv8(double d0, double d1, int g, int t) {
    const unsigned FULL = 0xffffffffu;
    double y0 = (g == 2 * t) ? 1.0 : 0.0;
    double y1 = (g == 2 * t + 1) ? 1.0 : 0.0;
    int bad = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int jt = j >> 1;
        const double dj = (j & 1) ? d1 : d0;
        const double piv = __shfl_sync(FULL, dj, j * 4 + jt);
        if (!(piv > 0.0)) bad = 1;
        const double r = rsqrt(piv);
        double lij = __shfl_sync(FULL, dj, g * 4 + jt) * r;                 // L[g][j] for rows below the pivot
        lij = (g > j) ? lij : ((g == j) ? piv * r : 0.0);
        const double lm0 = __shfl_sync(FULL, lij, (2 * t) * 4);             // L[2t][j], L[2t+1][j]
        const double lm1 = __shfl_sync(FULL, lij, (2 * t + 1) * 4);
        if (2 * t > j) d0 = fma(-lij, lm0, d0);
        if (2 * t + 1 > j) d1 = fma(-lij, lm1, d1);
        if (t == jt) {
            if (j & 1) d1 = lij; else d0 = lij;
        }
        // L Y = I by forward elimination: row j of Y is final after scaling, rows below lose their multiple of it
        const double yj0 = __shfl_sync(FULL, y0, j * 4 + t) * r;
        const double yj1 = __shfl_sync(FULL, y1, j * 4 + t) * r;
        if (g == j) {
            y0 = yj0;
            y1 = yj1;
        } else if (g > j) {
            y0 = fma(-lij, yj0, y0);
            y1 = fma(-lij, yj1, y1);
        }
    }
    CfChol8 o;
    o.d0 = d0; o.d1 = d1; o.y0 = y0; o.y1 = y1; o.bad = bad;
    return o;
}

// acc (64x64 SPD tile, lower part valid) <- its Cholesky factor; dp <- packed blocks + diagonal-block inverses.
// dg0 / dg1: this lane's diagonal entry of L in the warp's first / second row tile (1.0 for lanes that hold none).
__device__ __forceinline__ void cf_potrf_regs(CfAcc& acc, double* dp, int warp, int g, int t, bool& bad, double& dg0, double& dg1) {
    const int off = g * 8 + 2 * t;
#pragma unroll
    for (int p = 0; p < 8; p++) {
        if (warp == (p >> 1)) {
            const CfChol8 c8 = cf_chol8_inv8(acc[p & 1][p][0], acc[p & 1][p][1], g, t);
            acc[p & 1][p][0] = c8.d0;
            acc[p & 1][p][1] = c8.d1;
            bad = bad || c8.bad;
            const double dv = (g == 2 * t) ? c8.d0 : ((g == 2 * t + 1) ? c8.d1 : 1.0);
            if (p & 1) dg1 = dv; else dg0 = dv;
            *reinterpret_cast<double2*>(dp + dp_block(p, p) + off) = make_double2(c8.d0, c8.d1);
            *reinterpret_cast<double2*>(dp + dp_inv(p) + off) = make_double2(c8.y0, c8.y1);
        }
        if (p == 7) break;
        cf_consumer_bar();
        const double2 bi = *reinterpret_cast<const double2*>(dp + dp_inv(p) + off);
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            const int m = 2 * warp + mt;
            if (m > p) {
                double x0 = 0.0, x1 = 0.0;
                dmma884(x0, x1, acc[mt][p][0], bi.x);
                dmma884(x0, x1, acc[mt][p][1], bi.y);
                acc[mt][p][0] = x0;
                acc[mt][p][1] = x1;
                *reinterpret_cast<double2*>(dp + dp_block(m, p) + off) = make_double2(x0, x1);
            }
        }
        cf_consumer_bar();
#pragma unroll
        for (int q = p + 1; q < 8; q++) {
            if (2 * warp + 1 < q) continue;
            const double2 bl = *reinterpret_cast<const double2*>(dp + dp_block(q, p) + off);
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
                if (2 * warp + mt >= q) {
                    dmma884(acc[mt][q][0], acc[mt][q][1], -acc[mt][p][0], bl.x);
                    dmma884(acc[mt][q][0], acc[mt][q][1], -acc[mt][p][1], bl.y);
                }
            }
        }
    }
}

// y = L_kk^-1 rhs for one 64-row block, by one warp, from the packed diagonal block (36 lower 8x8 blocks + the inverses of
// the 8 diagonal ones): eight dependent 8x8 steps  v = rhs_p - sum_{q<p} L_pq y_q,  y_p = inv(L_pp) v.  Lane (g, t) works on
// row g and the column pair (2t, 2t+1) of every 8x8 block; sums over t by quad shuffles.  Result: yout[8 p + g].
// One copy of the code (noinline, like cf_chol8_inv8): its registers stay out of the kernel's GEMM loops.  tk: the block's
// entries of the right-hand side t; rhs: -(L_k,0..k-1 y) collected by the GEMM.
__device__ __noinline__ void cf_forward_block(const double* dp, const double* rhs, const double* tk, double* yout, int g, int t) {
    const unsigned FULL = 0xffffffffu;
    const int off = g * 8 + 2 * t;
    double tv[8];
#pragma unroll
    for (int p = 0; p < 8; p++) tv[p] = tk ? tk[p * 8 + g] : 0.0;
    double y0[8], y1[8];      // y_q[2t], y_q[2t+1] of the finished blocks
#pragma unroll
    for (int p = 0; p < 8; p++) {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < p; q++) {
            const double2 l = *reinterpret_cast<const double2*>(dp + dp_block(p, q) + off);
            acc = fma(l.x, y0[q], acc);
            acc = fma(l.y, y1[q], acc);
        }
        acc += __shfl_xor_sync(FULL, acc, 1);
        acc += __shfl_xor_sync(FULL, acc, 2);
        const double v = (tv[p] + rhs[p * 8 + g]) - acc;                // v[g], the same in the four lanes of a quad
        const double v0 = __shfl_sync(FULL, v, (2 * t) * 4), v1 = __shfl_sync(FULL, v, (2 * t + 1) * 4);
        const double2 iv = *reinterpret_cast<const double2*>(dp + dp_inv(p) + off);
        double yp = fma(iv.x, v0, iv.y * v1);
        yp += __shfl_xor_sync(FULL, yp, 1);
        yp += __shfl_xor_sync(FULL, yp, 2);                            // y_p[g]
        y0[p] = __shfl_sync(FULL, yp, (2 * t) * 4);
        y1[p] = __shfl_sync(FULL, yp, (2 * t + 1) * 4);
        if (t == 0) yout[p * 8 + g] = yp;
    }
}

// Anti-transposed copy of a finished tile: element (R, C) of L' goes to vt[np-1-C][np-1-R].  For a fixed (mt, nt, j) the 8
// lanes g of a quad column write 8 consecutive doubles (64 bytes) of one row; plain stores, nobody in this launch reads them.
__device__ __forceinline__ void cf_store_antitransposed(const CfAcc& acc, double* vt, int np, int ldd, int row0, int col0, int warp,
                                                        int g, int t, bool diag) {
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
        const int R = row0 + warp * 16 + mt * 8 + g;
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            const bool lower = !diag || nt <= 2 * warp + mt;
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int C = col0 + nt * 8 + 2 * t + j;
                vt[(size_t)(np - 1 - C) * ldd + (np - 1 - R)] = lower ? acc[mt][nt][j] : 0.0;
            }
        }
    }
}

// ---- the kernel -----------------------------------------------------------------------------------------------------
// SYRK: the source is the fused M' = P (I + L_K^T W L_K) P (two instantiations keep the plain path free of its registers)
// FWD: fused forward substitution (p.fwd_*); separate instantiations keep the plain factorisation free of its registers.
// MINCTAS: resident CTAs per SM the kernel is compiled for.  3 (128 registers, a few spills) keeps the tensor pipe busiest when
// every CTA slot has independent tasks (>= ~200 chains); 2 (168 registers, no spills, grid 2 x SMs) runs every task faster and
// wins when the launch is bound by the chains' critical paths (measured, 64 / 141 / 256 / 512 chains: k_chol 3.30 / 5.78 / 9.68 /
// 18.69 ms per FULL estimate against 3.61 / 5.94 / 9.49 / 17.81).  Same arithmetic per task: the host picks by the batch size.
template <bool SYRK, bool FWD = false, int MINCTAS = APM_CF_MIN_CTAS>
__global__ void __launch_bounds__(CF_THREADS, MINCTAS) k_chol_flow(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tms, const __grid_constant__ CUtensorMap tml,
                                                                           CholFlowParams p) {
    extern __shared__ unsigned char cf_smem_raw[];
    const uint32_t raw = smem_u32(cf_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = cf_smem_raw + (base - raw);
    unsigned char* ring = sm;                                                  // stage s: A box at s*16 KB, B box 8 KB behind
    double* dp = reinterpret_cast<double*>(sm + CF_STAGES * 2 * CF_CHUNK_BYTES);
    unsigned char* ctrl = reinterpret_cast<unsigned char*>(dp) + DP_BYTES;
    const uint32_t ring_u = base, dp_u = base + CF_STAGES * 2 * CF_CHUNK_BYTES, ctrl_u = dp_u + DP_BYTES;
    // control block: mbarriers (8 B each) then task descriptors and scratch
    const uint32_t bar_full = ctrl_u, bar_empty = ctrl_u + 8 * CF_STAGES;
    const uint32_t bar_tq_full = ctrl_u + 16 * CF_STAGES, bar_tq_empty = bar_tq_full + 16;
    const uint32_t bar_dp_full = bar_tq_empty + 16, bar_dp_empty = bar_dp_full + 8;
    volatile CfTask* tq = reinterpret_cast<volatile CfTask*>(ctrl + 16 * CF_STAGES + 48);       // 2 descriptors
    double* red = reinterpret_cast<double*>(ctrl + 16 * CF_STAGES + 48 + 2 * sizeof(CfTask));   // 4 partial log-dets
    double* wst = reinterpret_cast<double*>(ctrl + CF_WST_OFF);                                         // W_r of a TN stage: [stage][16]
    const uint32_t wst_u = ctrl_u + CF_WST_OFF;
    double* bst = reinterpret_cast<double*>(ctrl + CF_BST_OFF);                                        // b_r of a TN stage: [stage][16]
    const uint32_t bst_u = ctrl_u + CF_BST_OFF;
    double* fw_rhs = reinterpret_cast<double*>(ctrl + CF_RHS_OFF);                                     // [2][64]: right-hand side block of a diag task (by task parity)
    static_assert(16 * CF_STAGES + 48 + 2 * sizeof(CfTask) + 4 * 8 <= CF_WST_OFF, "control block too small");

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < CF_STAGES; s++) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, CF_CONSUMER_WARPS);
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(bar_tq_full + 8 * s, 1);
            mbar_init(bar_tq_empty + 8 * s, CF_CONSUMER_WARPS);
        }
        mbar_init(bar_dp_full, 1);
        mbar_init(bar_dp_empty, CF_CONSUMER_WARPS);
        fence_mbar_init();
    }
    __syncthreads();
    const int nb = p.nb;

    if (warp == CF_CONSUMER_WARPS) {
        // ================================ producer ================================
        if (lane != 0) return;
        uint32_t it = 0;
        int n = 0;
        const int nact = p.counter[1];
        const int total_tasks = nact * (nb * (nb + 1) / 2);
        for (;;) {
            const int tix = atomicAdd(p.counter, 1);
            int type = -1, k = 0, b = 0, i = 0;
            if (tix < total_tasks) {
                // step-major order: step k = [diag(k) of every chain][panel(k, i) chain-major]
                int r = tix;
                while (r >= nact * (nb - k)) { r -= nact * (nb - k); k++; }
                if (r < nact) {
                    type = 1; b = p.list[r]; i = k;
                } else {
                    r -= nact;
                    type = 0; b = p.list[r / (nb - k - 1)]; i = k + 1 + r % (nb - k - 1);
                }
            }
            const int slot = n & 1;
            mbar_wait(bar_tq_empty + 8 * slot, ((n >> 1) & 1) ^ 1);
            tq[slot].type = type; tq[slot].k = k; tq[slot].b = b; tq[slot].i = i;
            mbar_arrive(bar_tq_full + 8 * slot);
            if (type < 0) break;
            const int m = p.dst_idx ? p.dst_idx[b] : p.dst_m0 + b;
            const int row0 = m * p.np;
            const int* prog = p.progress + (size_t)b * nb;
            if (!SYRK) {
                // source tile (i, k): four boxes in two stages.  No dependency: nobody writes this tile before this task does.
                const int srow = (p.src_idx ? p.src_idx[b] : p.src_m0 + b) * p.np + i * TB;
                for (int h = 0; h < 2; h++, it++) {
                    const uint32_t s = it % CF_STAGES;
                    mbar_wait(bar_empty + 8 * s, ((it / CF_STAGES) & 1) ^ 1);
                    mbar_expect_tx(bar_full + 8 * s, 2 * CF_CHUNK_BYTES);
                    tma_load_2d(ring_u + s * 2 * CF_CHUNK_BYTES, &tms, k * TB + 32 * h, srow, bar_full + 8 * s);
                    tma_load_2d(ring_u + s * 2 * CF_CHUNK_BYTES + CF_CHUNK_BYTES, &tms, k * TB + 32 * h + 16, srow, bar_full + 8 * s);
                }
            } else {
                // source tile of M' from L_K: 4 (k + 1) stages of 16 rows, column blocks I = nb-1-i (A) and K = nb-1-k (B)
                const int I = nb - 1 - i, K = nb - 1 - k;
                const int lrow = p.lk_idx[b] * p.np + K * TB;
                const double* wsrc = p.w + (long long)b * p.w_bs + K * TB;
                const bool dg = type != 0;
                for (int c = 0; c < 4 * (k + 1); c++, it++) {
                    const uint32_t s = it % CF_STAGES;
                    mbar_wait(bar_empty + 8 * s, ((it / CF_STAGES) & 1) ^ 1);
                    const bool lt = FWD && dg;
                    mbar_expect_tx(bar_full + 8 * s, (dg ? CF_CHUNK_BYTES : 2 * CF_CHUNK_BYTES) + 128 + (lt ? 128 : 0));
                    const uint32_t st = ring_u + s * 2 * CF_CHUNK_BYTES;
                    for (int j = 0; j < 4; j++) {
                        tma_load_2d(st + CF_CHUNK_BYTES + j * 2048, &tml, K * TB + 16 * j, lrow + 16 * c, bar_full + 8 * s);
                        if (!dg) tma_load_2d(st + j * 2048, &tml, I * TB + 16 * j, lrow + 16 * c, bar_full + 8 * s);
                    }
                    bulk_load_1d(wst_u + s * 128, wsrc + 16 * c, 128, bar_full + 8 * s);
                    if (lt) bulk_load_1d(bst_u + s * 128, p.lt_b + (long long)b * p.fwd_bs + K * TB + 16 * c, 128, bar_full + 8 * s);
                }
            }
            if (type == 0) {
                if (k > 0) {
                    while (cf_ld_relaxed(prog + i) < k) __nanosleep(p.spin_ns);
                    while (cf_ld_relaxed(prog + k) < k) __nanosleep(p.spin_ns);
                    __threadfence();
                    fence_proxy_async_global();
                    for (int c = 0; c < 4 * k; c++, it++) {
                        const uint32_t s = it % CF_STAGES;
                        mbar_wait(bar_empty + 8 * s, ((it / CF_STAGES) & 1) ^ 1);
                        mbar_expect_tx(bar_full + 8 * s, 2 * CF_CHUNK_BYTES);
                        tma_load_2d(ring_u + s * 2 * CF_CHUNK_BYTES, &tm, c * CF_KC, row0 + i * TB, bar_full + 8 * s);
                        tma_load_2d(ring_u + s * 2 * CF_CHUNK_BYTES + CF_CHUNK_BYTES, &tm, c * CF_KC, row0 + k * TB, bar_full + 8 * s);
                    }
                }
                while (cf_ld_relaxed(prog + k) < k + 1) __nanosleep(p.spin_ns);
                __threadfence();
                fence_proxy_async_global();
                if (n > 0) mbar_wait(bar_dp_empty, (n - 1) & 1);
                mbar_expect_tx(bar_dp_full, DP_BYTES);
                bulk_load_1d(dp_u, p.diagpack + ((size_t)b * nb + k) * DP_DOUBLES, DP_BYTES, bar_dp_full);
            } else {
                int known = 0, yknown = 0;
                const bool fwd = FWD;
                const double* ysrc = fwd ? p.fwd_y + (long long)b * p.fwd_bs : nullptr;
                for (int j = 0; j < k; j++) {
                    if (known < j + 1 || (fwd && yknown < j + 1)) {
                        while ((known = cf_ld_relaxed(prog + k)) < j + 1) __nanosleep(p.spin_ns);
                        if (fwd)
                            while ((yknown = cf_ld_relaxed(p.yprog + b)) < j + 1) __nanosleep(p.spin_ns);
                        __threadfence();
                        fence_proxy_async_global();
                    }
                    for (int c = 4 * j; c < 4 * j + 4; c++, it++) {
                        const uint32_t s = it % CF_STAGES;
                        mbar_wait(bar_empty + 8 * s, ((it / CF_STAGES) & 1) ^ 1);
                        mbar_expect_tx(bar_full + 8 * s, CF_CHUNK_BYTES + (fwd ? 128 : 0));
                        tma_load_2d(ring_u + s * 2 * CF_CHUNK_BYTES, &tm, c * CF_KC, row0 + k * TB, bar_full + 8 * s);
                        if (fwd) bulk_load_1d(wst_u + s * 128, ysrc + 16 * c, 128, bar_full + 8 * s);   // y entries of the chunk's columns
                    }
                }
                if (n > 0) mbar_wait(bar_dp_empty, (n - 1) & 1);   // keeps the producer within one task of the consumers
            }
            n++;
        }
        return;
    }

    // ================================ consumers ================================
    const int g = lane >> 2, t = lane & 3;
    uint32_t it = 0;
    int n = 0, pc = 0;
    for (;; n++) {
        const int slot = n & 1;
        mbar_wait(bar_tq_full + 8 * slot, (n >> 1) & 1);
        const int type = tq[slot].type, k = tq[slot].k, b = tq[slot].b, i = tq[slot].i;
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tq_empty + 8 * slot);
        if (type < 0) break;
        double* dst = p.dst + chain_index(p.dst_idx, b) * p.dst_bs;
        const double* sc = p.scale ? p.scale + (long long)b * p.scale_bs : nullptr;
        int* prog = p.progress + (size_t)b * nb;
        const bool diag = type != 0;
        CfAcc acc;
        double tq[2] = {0.0, 0.0};       // <true, true>: this lane's share of t'[64 k + 16 warp + 8 mt + g]
        if (!SYRK) {
            // ---- source tile A_ik (two stages of the ring)
            const double* rs = sc ? sc + i * TB : nullptr;
            const double* cs = sc ? sc + k * TB : nullptr;
            const bool ident = diag && p.add_identity != 0;
#pragma unroll
            for (int h = 0; h < 2; h++, it++) {
                const uint32_t s = it % CF_STAGES;
                mbar_wait(bar_full + 8 * s, (it / CF_STAGES) & 1);
                cf_src_from_stage(acc, ring + s * 2 * CF_CHUNK_BYTES, h, rs, cs, ident, warp, g, t);
                uint32_t dep = 0;
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                    for (int q = 0; q < 4; q++) dep |= (uint32_t)__double2hiint(acc[mt][4 * h + q][1]);
                cf_release_stage(bar_empty + 8 * s, dep, p.zero, lane);
            }
        } else {
            // ---- source tile of M' = P (I + L_K^T W L_K) P accumulated from L_K (never stored)
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int nt = 0; nt < 8; nt++) {
                    const bool dtile = diag && nt == 2 * warp + mt;
                    acc[mt][nt][0] = (dtile && g == 2 * t) ? 1.0 : 0.0;
                    acc[mt][nt][1] = (dtile && g == 2 * t + 1) ? 1.0 : 0.0;
                }
            for (int c = 0; c < 4 * (k + 1); c++, it++) {
                const uint32_t s = it % CF_STAGES;
                mbar_wait(bar_full + 8 * s, (it / CF_STAGES) & 1);
                const unsigned char* st = ring + s * 2 * CF_CHUNK_BYTES;
                uint32_t dep = 0;
                if (diag) {
                    cf_gemm_chunk_tn<true, FWD>(acc, st + CF_CHUNK_BYTES, st + CF_CHUNK_BYTES, wst + s * 16, warp, g, t, bst + s * 16, tq);
                    if (FWD) dep = (uint32_t)__double2hiint(tq[0]) | (uint32_t)__double2hiint(tq[1]);
                } else {
                    cf_gemm_chunk_tn<false>(acc, st, st + CF_CHUNK_BYTES, wst + s * 16, warp, g, t);
                }
                cf_release_stage(bar_empty + 8 * s, cf_acc_dep(acc) | dep, p.zero, lane);
            }
        }
        if (!diag) {
            // ---- panel(k, i): T = A_ik - L_i,0..k-1 L_k,0..k-1^T ; L_ik = T L_kk^-T
            for (int c = 0; c < 4 * k; c++, it++) {
                const uint32_t s = it % CF_STAGES;
                mbar_wait(bar_full + 8 * s, (it / CF_STAGES) & 1);
                cf_gemm_chunk<false>(acc, ring + s * 2 * CF_CHUNK_BYTES, ring + s * 2 * CF_CHUNK_BYTES + CF_CHUNK_BYTES, warp, g, t);
                cf_release_stage(bar_empty + 8 * s, cf_acc_dep(acc), p.zero, lane);
            }
            mbar_wait(bar_dp_full, pc & 1);
            pc++;
            cf_trsm_regs<false>(acc, dp, warp, g, t);
            double* out = dst + (size_t)i * TB * p.ldd + k * TB;
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int nt = 0; nt < 8; nt++)
                    *reinterpret_cast<double2*>(out + (size_t)(warp * 16 + mt * 8 + g) * p.ldd + nt * 8 + 2 * t) =
                        make_double2(acc[mt][nt][0], acc[mt][nt][1]);
            fence_proxy_async_global();             // this thread's tile stores (generic proxy) will be read by other CTAs' TMA
            // every warp's stores precede warp 0's release (cumulativity through the barrier); warps 1..3 do not wait
            if (warp == 0) {
                cf_publish_sync(n & 1);
                if (lane == 0) cf_st_release(prog + i, k + 1);
            } else {
                cf_publish_arrive(n & 1);
            }
            if (SYRK && p.vt_out)
                cf_store_antitransposed(acc, p.vt_out + (long long)p.lk_idx[b] * p.vt_bs, p.np, p.ldd, i * TB, k * TB, warp, g, t, false);
        } else {
            // ---- diag(k): D = A_kk - L_k,0..k-1 L_k,0..k-1^T (tiles on / below the diagonal) ; L_kk = chol(D)
            const bool fwd = FWD;
            if (fwd) {
                // fused forward substitution: rhs_k = t_k - sum_j L_kj y_j, collected from the A fragments of the GEMM
                double ys[2] = {0.0, 0.0};
                for (int c = 0; c < 4 * k; c++, it++) {
                    const uint32_t s = it % CF_STAGES;
                    mbar_wait(bar_full + 8 * s, (it / CF_STAGES) & 1);
                    cf_gemm_chunk<true, true>(acc, ring + s * 2 * CF_CHUNK_BYTES, ring + s * 2 * CF_CHUNK_BYTES, warp, g, t, wst + s * 16, ys);
                    cf_release_stage(bar_empty + 8 * s, cf_acc_dep(acc) | (uint32_t)__double2hiint(ys[0]) | (uint32_t)__double2hiint(ys[1]), p.zero, lane);
                }
#pragma unroll
                for (int mt = 0; mt < 2; mt++) {
                    if (SYRK) ys[mt] += tq[mt];      // right-hand side entries accumulated with M' (same index as the row)
                    ys[mt] += __shfl_xor_sync(0xffffffffu, ys[mt], 1);
                    ys[mt] += __shfl_xor_sync(0xffffffffu, ys[mt], 2);
                }
                if (t == 0) {
                    fw_rhs[(n & 1) * 64 + warp * 16 + g] = ys[0];
                    fw_rhs[(n & 1) * 64 + warp * 16 + 8 + g] = ys[1];
                }
            } else {
                for (int c = 0; c < 4 * k; c++, it++) {
                    const uint32_t s = it % CF_STAGES;
                    mbar_wait(bar_full + 8 * s, (it / CF_STAGES) & 1);
                    cf_gemm_chunk<true>(acc, ring + s * 2 * CF_CHUNK_BYTES, ring + s * 2 * CF_CHUNK_BYTES, warp, g, t);
                    cf_release_stage(bar_empty + 8 * s, cf_acc_dep(acc), p.zero, lane);
                }
            }
            bool bad = false;
            double dg0 = 1.0, dg1 = 1.0;
            cf_consumer_bar();      // the other warps may still read dp for the previous task (triangular solve / inverse block)
            cf_potrf_regs(acc, dp, warp, g, t, bad, dg0, dg1);
            // L_kk (explicit zeros above the diagonal) -> dst
            double* out = dst + (size_t)k * TB * p.ldd + k * TB;
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int nt = 0; nt < 8; nt++) {
                    const bool lower = nt <= 2 * warp + mt;
                    *reinterpret_cast<double2*>(out + (size_t)(warp * 16 + mt * 8 + g) * p.ldd + nt * 8 + 2 * t) =
                        make_double2(lower ? acc[mt][nt][0] : 0.0, lower ? acc[mt][nt][1] : 0.0);
                }
            // partial log-det from the diagonal entries kept by cf_potrf_regs
            double lg = (g >> 1 == t) ? log(dg0) + log(dg1) : 0.0;
            lg = warp_sum(lg);
            if (lane == 0) red[warp] = lg;
            if (__any_sync(0xffffffffu, bad) && lane == 0) atomicMax(&p.status[b], p.fail_code);
            fence_proxy_async_smem();               // this thread's writes to dp (generic proxy) before the bulk store reads them
            fence_proxy_async_global();             // ... and its stores of the L_kk tile before other CTAs' TMA reads
            cf_consumer_bar();
            if (tid == 0) {
                bulk_store_1d(p.diagpack + ((size_t)b * nb + k) * DP_DOUBLES, dp_u, DP_BYTES);
                if (p.logdet_parts)
                    p.logdet_parts[(size_t)chain_index(p.logdet_idx, b) * p.logdet_stride + k] = (red[0] + red[1]) + (red[2] + red[3]);
                bulk_store_wait();
                fence_proxy_async_global();
                cf_st_release(prog + k, k + 1);
            }
            if (SYRK && p.vt_out)
                cf_store_antitransposed(acc, p.vt_out + (long long)p.lk_idx[b] * p.vt_bs, p.np, p.ldd, k * TB, k * TB, warp, g, t, true);
            if (p.inv_out) {
                // (L_kk^{-1})^T = I * L_kk^{-T} for the single right-hand-side solves of the Newton step (k_trsv2)
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                    for (int nt = 0; nt < 8; nt++) {
                        const bool dtile = nt == 2 * warp + mt;
                        acc[mt][nt][0] = (dtile && g == 2 * t) ? 1.0 : 0.0;
                        acc[mt][nt][1] = (dtile && g == 2 * t + 1) ? 1.0 : 0.0;
                    }
                cf_trsm_regs<true>(acc, dp, warp, g, t);
                double* io = p.inv_out + (long long)b * p.inv_bs + (size_t)k * TB * TB;
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                    for (int nt = 0; nt < 8; nt++)
                        *reinterpret_cast<double2*>(io + (warp * 16 + mt * 8 + g) * TB + nt * 8 + 2 * t) =
                            make_double2(acc[mt][nt][0], acc[mt][nt][1]);
            }
            if (fwd && warp == 0) {
                // y_k = L_kk^-1 rhs_k: off the factorisation's critical path (L_kk is already published); the next diagonal
                // task of this chain waits for yprog before it streams y_k
                cf_forward_block(dp, fw_rhs + (n & 1) * 64, SYRK ? nullptr : p.fwd_t + (long long)b * p.fwd_bs + k * TB,
                                 p.fwd_y + (long long)b * p.fwd_bs + k * TB, g, t);
                fence_proxy_async_global();         // read by other CTAs' bulk copies
                __syncwarp();
                if (lane == 0) cf_st_release(p.yprog + b, k + 1);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_dp_empty);   // dp may be overwritten by the next panel task's bulk copy
    }
}

}  // namespace apm
