// common.cuh -- shared device helpers for the apm_b200 kernels (sm_100a, fp64).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace apm {

constexpr int TB = 64;          // tile edge of every blocked kernel; matrices are padded to a multiple
#ifndef APM_KC
#define APM_KC 16
#endif
#ifndef APM_STAGES
#define APM_STAGES 3
#endif
#ifndef APM_MIN_CTAS
#define APM_MIN_CTAS 3
#endif
constexpr int KC = APM_KC;      // k-chunk (doubles) staged per pipeline stage
#ifndef APM_SWIZZLE
#define APM_SWIZZLE 1
#endif
// smem row stride of a k-chunk: padded by 4 doubles ((KC+4)*8 B == 32 mod 128), or unpadded with the 16-byte
// segments of a row XOR-swizzled by ((row & 3) << 1) -- both give conflict-free DMMA fragment loads
constexpr bool SWIZZLE = APM_SWIZZLE != 0;
constexpr int KCP = SWIZZLE ? KC : KC + 4;
constexpr int STAGES = APM_STAGES;   // cp.async pipeline depth
constexpr int MIN_CTAS = APM_MIN_CTAS;
constexpr int TILE_THREADS = 128;
constexpr int TSP = 68;         // row stride of the 64x64 fp64 work tile: 68 == 4 (mod 16) -> conflict-free DMMA fragment loads
constexpr int VSP = 65;         // row stride of tiles accessed one row / one column per thread (vec kernels)
constexpr int GEMM_SMEM_DOUBLES = 2 * STAGES * TB * KCP;             // A and B stages
constexpr int TILE_SCRATCH_DOUBLES = TB * TSP + 36 * 64 + TB + 8;      // work tile + packed diagonal block + reciprocals + flags
constexpr int TILE_SMEM_BYTES = (GEMM_SMEM_DOUBLES > TILE_SCRATCH_DOUBLES ? GEMM_SMEM_DOUBLES : TILE_SCRATCH_DOUBLES) * 8;  // 70208 B -> 3 CTAs / SM

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// One fp64 tensor-core instruction: D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4.
// Fragment layout (g = lane/4, t = lane%4): a = A[g][t], b = B[k=t][n=g], c0/c1 = C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

// log of the standard normal CDF with the branch structure of scipy.special.log_ndtr (scipy 1.18):
//   x < -1 : log(erfcx(-x/sqrt2)/2) - x^2/2   else   log1p(-erfc(x/sqrt2)/2)
// (call sites in the reference: lpa.py:86, 105; estimators.py:229, 324)
__device__ __forceinline__ double log_ndtr(double x) {
    const double t = x * 0.70710678118654752440;
    if (x < -1.0) return log(erfcx(-t) * 0.5) - t * t;
    return log1p(-erfc(t) * 0.5);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// pointer of chain b inside a batched buffer, with optional slot indirection
__device__ __forceinline__ long long chain_index(const int* idx, int b) { return idx ? (long long)idx[b] : (long long)b; }

}  // namespace apm
