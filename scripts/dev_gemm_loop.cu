// dev micro-benchmark: what bounds the 64x64 DMMA tile GEMM loop?  (build: see scripts/dev_gemm_loop.sh)
//   mode 0: LDS + DMMA only (operands resident in shared memory, no barriers, no global loads)
//   mode 1: mode 0 + one __syncthreads() per k-chunk
//   mode 2: the product micro-kernel gemm_nt_64x64 (cp.async pipeline), operands L2-resident
//   mode 3: as 2 with every CTA streaming its own DRAM-resident operands
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../auxiliary-pm-mcmc_b200/csrc/tile_engine.cuh"
using namespace apm;

template <int MODE>
__global__ void __launch_bounds__(TILE_THREADS, MIN_CTAS) k_loop(const double* A, const double* Bm, int ld, int kdepth, double* out,
                                                                long long cta_stride) {
    extern __shared__ __align__(16) double smem[];
    Acc acc;
    acc.zero();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3, wm = warp >> 1, wn = warp & 1;
    if (MODE >= 2) {
        const double* a = A + (MODE == 3 ? (long long)blockIdx.x * cta_stride : 0);
        const double* b = Bm + (MODE == 3 ? (long long)blockIdx.x * cta_stride : 0);
        gemm_nt_64x64<true>(acc, a, ld, b, ld, kdepth, smem);
    } else {
        double* smA = smem;
        double* smB = smem + STAGES * TB * KCP;
        for (int e = tid; e < 2 * STAGES * TB * KCP; e += TILE_THREADS) smem[e] = 1e-3 * (e % 97);
        __syncthreads();
        const int nchunks = kdepth / KC;
        for (int kc = 0; kc < nchunks; kc++) {
            if (MODE == 1) __syncthreads();
            const int st = kc % STAGES;
            const double* a_s = smA + st * TB * KCP + (wm * 32 + g) * KCP + (SWIZZLE ? (t & 1) : t);
            const double* b_s = smB + st * TB * KCP + (wn * 32 + g) * KCP + (SWIZZLE ? (t & 1) : t);
            const int swz = (g & 3) << 1, th = t >> 1;
#pragma unroll
            for (int kk = 0; kk < KC / 4; kk++) {
                double a[4], b[4];
                const int ko = SWIZZLE ? (((kk * 2 + th) ^ swz) * 2) : kk * 4;
#pragma unroll
                for (int mi = 0; mi < 4; mi++) a[mi] = -a_s[mi * 8 * KCP + ko];
#pragma unroll
                for (int ni = 0; ni < 4; ni++) b[ni] = b_s[ni * 8 * KCP + ko];
#pragma unroll
                for (int mi = 0; mi < 4; mi++)
#pragma unroll
                    for (int ni = 0; ni < 4; ni++) dmma884(acc.v[mi][ni][0], acc.v[mi][ni][1], a[mi], b[ni]);
            }
        }
    }
    double s = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) s += acc.v[i][j][0] + acc.v[i][j][1];
    if (s == 123.456) out[0] = s;
}

int main(int argc, char** argv) {
    const int kdepth = 64 * 64;
    const int ld = kdepth;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int max_ctas = sms * 3;
    double *A, *out;
    const long long cta_stride = 64LL * ld;
    cudaMalloc(&A, sizeof(double) * cta_stride * (max_ctas + 1) * 2);
    cudaMemset(A, 0, sizeof(double) * cta_stride * (max_ctas + 1) * 2);
    cudaMalloc(&out, 64);
    double* Bm = A + cta_stride * (max_ctas + 1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto run = [&](int mode, int occ) {
        const int smem = occ == 3 ? TILE_SMEM_BYTES : (occ == 2 ? 110 * 1024 : 200 * 1024);
        const int grid = sms * occ;
        void (*kern)(const double*, const double*, int, int, double*, long long) =
            mode == 0 ? k_loop<0> : mode == 1 ? k_loop<1> : mode == 2 ? k_loop<2> : k_loop<3>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            kern<<<grid, TILE_THREADS, smem>>>(A, Bm, ld, kdepth, out, cta_stride);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        const double flop = (double)grid * 2.0 * 64 * 64 * kdepth;
        printf("mode %d  %d CTA/SM: %.3f ms  %.2f TFLOP/s  (%s)\n", mode, occ, best, flop / best / 1e9, cudaGetErrorString(cudaGetLastError()));
    };
    for (int mode = 0; mode < 4; mode++)
        for (int occ = 1; occ <= 3; occ++) run(mode, occ);
    return 0;
}
