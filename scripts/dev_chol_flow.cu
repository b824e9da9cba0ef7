// dev_chol_flow.cu -- standalone check + timing of k_chol_flow (TMA / mbarrier dataflow Cholesky) against the round-1
// k_chol_dataflow and a host reconstruction.  Build: see scripts/build_dev.sh.  Usage: dev_chol_flow [n] [B] [reps] [grid_mult]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include "r1_chol_engine.cuh"
#include "../auxiliary-pm-mcmc_b200/csrc/chol_flow.cuh"
#include "../auxiliary-pm-mcmc_b200/csrc/tmap_host.h"

using namespace apm;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__global__ void k_make_K(double* K, const double* X, int n, int np, int D, const double* par) {
    const int b = blockIdx.z;
    const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x;
    if (i >= np || j >= np) return;
    double v;
    if (i >= n || j >= n) v = (i == j) ? 1.0 : 0.0;
    else {
        double acc = 0;
        for (int k = 0; k < D; k++) { const double d = (X[i * D + k] - X[j * D + k]) / par[b * 2 + 1]; acc += d * d; }
        v = par[b * 2] * exp(-0.5 * acc) + (i == j ? 1e-8 : 0.0);
    }
    K[(size_t)b * np * np + (size_t)i * np + j] = v;
}

static double max_rel_lower(const std::vector<double>& a, const std::vector<double>& b, int np, int B, double* worst_abs) {
    double mx = 0, ref = 0;
    for (int c = 0; c < B; c++)
        for (int i = 0; i < np; i++)
            for (int j = 0; j <= i; j++) {
                const size_t o = (size_t)c * np * np + (size_t)i * np + j;
                mx = std::max(mx, fabs(a[o] - b[o]));
                ref = std::max(ref, fabs(b[o]));
                if (a[o] != a[o]) mx = 1e300;
            }
    *worst_abs = mx;
    return mx / ref;
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 768;
    const int B = argc > 2 ? atoi(argv[2]) : 256;
    const int reps = argc > 3 ? atoi(argv[3]) : 5;
    const int grid_mult = argc > 4 ? atoi(argv[4]) : 0;
    const int D = 8;
    const int np = (n + 63) / 64 * 64, nb = np / 64;
    const size_t mat = (size_t)np * np;
    printf("n=%d np=%d nb=%d B=%d reps=%d  CF_SMEM=%d stages=%d\n", n, np, nb, B, reps, CF_SMEM_BYTES, CF_STAGES);
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));

    std::vector<double> hX((size_t)n * D), hpar(2 * B), hscale((size_t)B * np);
    srand(1);
    for (auto& v : hX) { double s = 0; for (int q = 0; q < 12; q++) s += rand() / (double)RAND_MAX; v = s - 6.0; }
    for (int b = 0; b < B; b++) { hpar[2 * b] = exp(0.5 + 0.3 * (rand() / (double)RAND_MAX - 0.5)); hpar[2 * b + 1] = 2.0 * exp(0.3 * (rand() / (double)RAND_MAX - 0.5)); }
    for (auto& v : hscale) v = 0.3 + 0.6 * rand() / (double)RAND_MAX;
    double *dX, *dpar, *dK, *dL0, *dL1, *dLK, *dscale, *dinv0, *dinv1, *dld0, *dld1, *ddp;
    int* dident;
    int *dstatus, *dactive, *dcounter, *dprogress, *dskip;
    CK(cudaMalloc(&dX, hX.size() * 8)); CK(cudaMalloc(&dpar, hpar.size() * 8)); CK(cudaMalloc(&dscale, hscale.size() * 8));
    CK(cudaMalloc(&dK, B * mat * 8)); CK(cudaMalloc(&dL0, B * mat * 8)); CK(cudaMalloc(&dL1, B * mat * 8)); CK(cudaMalloc(&dLK, B * mat * 8));
    CK(cudaMalloc(&dident, B * 4));
    { std::vector<int> id(B); for (int b = 0; b < B; b++) id[b] = b; CK(cudaMemcpy(dident, id.data(), B * 4, cudaMemcpyHostToDevice)); }
    CK(cudaMalloc(&dinv0, (size_t)B * nb * 4096 * 8)); CK(cudaMalloc(&dinv1, (size_t)B * nb * 4096 * 8));
    CK(cudaMalloc(&dld0, (size_t)B * nb * 8)); CK(cudaMalloc(&dld1, (size_t)B * nb * 8));
    CK(cudaMalloc(&ddp, (size_t)B * nb * DP_BYTES));
    CK(cudaMalloc(&dstatus, B * 4)); CK(cudaMalloc(&dactive, B * 4)); CK(cudaMalloc(&dcounter, 256)); CK(cudaMemset(dcounter, 0, 256)); CK(cudaMalloc(&dprogress, (size_t)B * nb * 4)); CK(cudaMalloc(&dskip, B * 4));
    CK(cudaMemcpy(dX, hX.data(), hX.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dpar, hpar.data(), hpar.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dscale, hscale.data(), hscale.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemset(dstatus, 0, B * 4));
    k_make_K<<<dim3((np + 15) / 16, (np + 15) / 16, B), dim3(16, 16)>>>(dK, dX, n, np, D, dpar);
    CK(cudaDeviceSynchronize());

    CK(cudaFuncSetAttribute(k_chol_dataflow, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM_BYTES));
    CK(cudaFuncSetAttribute(k_syrk_lk, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM_BYTES));
    CK(cudaFuncSetAttribute(k_chol_flow<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF_SMEM_BYTES));
    CK(cudaFuncSetAttribute(k_chol_flow<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF_SMEM_BYTES));
    int occ_old = 0, occ_new = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_old, k_chol_dataflow, TILE_THREADS, TILE_SMEM_BYTES));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_new, k_chol_flow<false>, CF_THREADS, CF_SMEM_BYTES));
    printf("occupancy: old %d CTAs/SM, new %d CTAs/SM, %d SMs\n", occ_old, occ_new, sms);
    CUtensorMap tmK, tm1, tmLK16;
    if (!make_matrix_tmap(&tm1, dL1, np, B) || !make_matrix_tmap(&tmK, dK, np, B) || !make_matrix_tmap(&tmLK16, dLK, np, B, 16)) { printf("tensor map creation failed\n"); return 2; }

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<double> h0(B * mat), h1(B * mat);

    for (int variant = 0; variant < 4; variant++) {
        // 0: chol(K) -> L;  1: chol(I + S K S) with inverse diagonal blocks;  2: as 1, in place, every 3rd chain inactive
        // 3: chol(M'), M' = P (I + L_K^T W L_K) P: round-1 k_syrk_lk + in-place Cholesky against the fused source of k_chol_flow
        const bool syrk = variant == 3;
        const bool scaled = variant == 1 || variant == 2, inplace = variant == 2;
        std::vector<int> hact(B, 1);
        if (variant == 2) for (int b = 0; b < B; b += 3) hact[b] = 0;
        CK(cudaMemcpy(dactive, hact.data(), B * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(dstatus, 0, B * 4));
        float ms_old = 0, ms_new = 0;
        // ---- round-1 kernel
        {
            CholParams p;
            p.src = (inplace || syrk) ? dL0 : dK; p.src_bs = (long long)mat; p.lds = np; p.src_idx = nullptr;
            p.dst = dL0; p.dst_bs = (long long)mat; p.ldd = np; p.dst_idx = nullptr;
            p.scale = scaled ? dscale : nullptr; p.scale_bs = np; p.add_identity = scaled;
            p.nb = nb; p.logdet_parts = dld0; p.logdet_stride = nb; p.logdet_idx = nullptr;
            p.inv_out = (scaled || syrk) ? dinv0 : nullptr; p.inv_bs = (long long)nb * 4096;
            p.status = dstatus; p.fail_code = 5; p.active = variant == 2 ? dactive : nullptr; p.nchains = B;
            p.sm_sem = nullptr; p.sem_limit = 0;
            CholFlow f;
            f.counter = dcounter; f.progress = dprogress; f.skip = dskip; f.group = B;
            f.total_tasks = B * (1 + nb * (nb - 1) / 2); f.flags = 0; f.spin_ns = 100;
            const int grid = std::min(occ_old * sms, f.total_tasks);
            for (int r = 0; r < reps + 1; r++) {
                if (inplace) CK(cudaMemcpy(dL0, dK, B * mat * 8, cudaMemcpyDeviceToDevice));
                if (r == 1) CK(cudaEventRecord(e0));
                CK(cudaMemsetAsync(dstatus, 0, B * 4));
                if (syrk) {
                    SyrkLkParams sp;
                    sp.LK = dLK; sp.lk_bs = (long long)mat; sp.ldk = np; sp.lk_idx = dident;
                    sp.W = dscale; sp.w_bs = np; sp.M = dL0; sp.m_bs = (long long)mat; sp.ldm = np;
                    sp.nb = nb; sp.ntiles = nb * (nb + 1) / 2; sp.status = dstatus; sp.mask = nullptr;
                    k_syrk_lk<<<B * sp.ntiles, TILE_THREADS, TILE_SMEM_BYTES>>>(sp);
                }
                CK(cudaMemsetAsync(dcounter, 0, 4));
                CK(cudaMemsetAsync(dprogress, 0, (size_t)B * nb * 4));
                k_chol_skip_snapshot<<<(B + 255) / 256, 256>>>(dstatus, p.active, dskip, B);
                k_chol_dataflow<<<grid, TILE_THREADS, TILE_SMEM_BYTES>>>(p, f);
                if (inplace && r >= 1) break;   // timing of in-place needs a fresh copy each time: time a single rep
            }
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
            CK(cudaEventElapsedTime(&ms_old, e0, e1));
            ms_old /= inplace ? 1 : reps;
        }
        // ---- new kernel
        {
            CholFlowParams p;
            p.src = inplace ? dL1 : dK; p.src_bs = (long long)mat;
            p.lk_idx = dident; p.w = dscale; p.w_bs = np; p.lds = np; p.src_idx = nullptr;
            p.dst = dL1; p.dst_bs = (long long)mat; p.ldd = np; p.dst_idx = nullptr; p.dst_m0 = 0; p.src_m0 = 0; p.zero = 0; p.np = np;
            p.scale = scaled ? dscale : nullptr; p.scale_bs = np; p.add_identity = scaled;
            p.nb = nb; p.logdet_parts = dld1; p.logdet_stride = nb; p.logdet_idx = nullptr;
            p.inv_out = (scaled || syrk) ? dinv1 : nullptr; p.inv_bs = (long long)nb * 4096;
            p.status = dstatus; p.fail_code = 5; p.active = variant == 2 ? dactive : nullptr; p.nchains = B;
            p.counter = dcounter + 4; p.progress = dprogress; p.list = dskip; p.diagpack = ddp; p.vt_out = nullptr; p.vt_bs = 0; p.fwd_t = nullptr; p.fwd_y = nullptr; p.fwd_bs = 0; p.yprog = nullptr; p.lt_b = nullptr;
            p.spin_ns = 64;
            const int total_tasks = B * nb * (nb + 1) / 2;
            int grid = std::min(occ_new * sms, total_tasks);
            if (grid_mult > 0) grid = std::min(grid, grid_mult * B);
            for (int r = 0; r < reps + 1; r++) {
                if (inplace) CK(cudaMemcpy(dL1, dK, B * mat * 8, cudaMemcpyDeviceToDevice));
                if (r == 1) CK(cudaEventRecord(e0));
                CK(cudaMemsetAsync(dstatus, 0, B * 4));
                k_chol_flow_init<<<(B * nb + 255) / 256, 256>>>(p.counter, p.progress, p.list, dstatus, p.active, B, nb, nullptr, nullptr);
                if (syrk) k_chol_flow<true, false, 3><<<grid, CF_THREADS, CF_SMEM_BYTES>>>(tm1, tmK, tmLK16, p);
                else k_chol_flow<false, false, 3><<<grid, CF_THREADS, CF_SMEM_BYTES>>>(tm1, inplace ? tm1 : tmK, tmLK16, p);
                if (inplace && r >= 1) break;
            }
            CK(cudaEventRecord(e1));
            cudaError_t es = cudaEventSynchronize(e1);
            if (es != cudaSuccess) { printf("new kernel failed: %s\n", cudaGetErrorString(es)); return 3; }
            CK(cudaGetLastError());
            CK(cudaEventElapsedTime(&ms_new, e0, e1));
            ms_new /= inplace ? 1 : reps;
#ifdef APM_CF_DBG_VERIFY
            {
                int hc[32];
                CK(cudaMemcpy(hc, dcounter, 128, cudaMemcpyDeviceToHost));
                printf("verify: (warp-chunk checks) stale A %d, stale B %d; first: k=%d i=%d c=%d b=%d code=%d\n", hc[12], hc[13], hc[15], hc[16], hc[17], hc[18], hc[19]);
                CK(cudaMemset(dcounter, 0, 256));
            }
#endif
        }
        if (variant == 0) CK(cudaMemcpy(dLK, dL1, B * mat * 8, cudaMemcpyDeviceToDevice));   // L_K of the M' variant
        CK(cudaMemcpy(h0.data(), dL0, B * mat * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h1.data(), dL1, B * mat * 8, cudaMemcpyDeviceToHost));
        if (variant == 2)   // inactive chains keep the copied source: identical in both
            for (int b = 0; b < B; b += 3) std::copy(h0.begin() + b * mat, h0.begin() + (b + 1) * mat, h1.begin() + b * mat);
        double wabs = 0;
        const double rel = max_rel_lower(h1, h0, np, B, &wabs);
        if (getenv("TILEMAP")) {
            int shown = 0;
            for (int c = 0; c < B && shown < 3; c++) {
                double cm = 0;
                std::vector<double> tm((size_t)nb * nb, 0.0);
                for (int i = 0; i < np; i++)
                    for (int j = 0; j <= i; j++) {
                        const size_t o = (size_t)c * mat + (size_t)i * np + j;
                        double d = fabs(h1[o] - h0[o]);
                        if (h1[o] != h1[o]) d = 9e99;
                        tm[(i / 64) * nb + j / 64] = std::max(tm[(i / 64) * nb + j / 64], d);
                        cm = std::max(cm, d);
                    }
                if (cm < 1e-9) continue;
                shown++;
                printf("chain %d: per-tile max |new - old| (rows i, cols k)\n", c);
                for (int i = 0; i < nb; i++) {
                    for (int k = 0; k <= i; k++) printf(" %8.1e", tm[i * nb + k]);
                    printf("\n");
                }
            }
        }
        // explicit zeros above the diagonal inside the diagonal tiles
        double upper = 0;
        for (int c = 0; c < B; c++) {
            if (variant == 2 && c % 3 == 0) continue;
            for (int i = 0; i < np; i++) for (int j = i + 1; j < (i / 64 + 1) * 64; j++) upper = std::max(upper, fabs(h1[c * mat + (size_t)i * np + j]));
        }
        // host reconstruction of chain 1: || L L^T - A ||_max / ||A||_max
        double rec = 0, an = 1;
        if (!syrk) {
            an = 0;
            const int c = 1 % B;
            std::vector<double> hK(mat);
            CK(cudaMemcpy(hK.data(), dK + c * mat, mat * 8, cudaMemcpyDeviceToHost));
            const double* L = h1.data() + c * mat;
            for (int i = 0; i < np; i += 7)
                for (int j = 0; j <= i; j += 3) {
                    double s = 0;
                    for (int k = 0; k <= j; k++) s += L[(size_t)i * np + k] * L[(size_t)j * np + k];
                    double a = hK[(size_t)i * np + j];
                    if (scaled) a = hscale[c * np + i] * a * hscale[c * np + j] + (i == j ? 1.0 : 0.0);
                    rec = std::max(rec, fabs(s - a));
                    an = std::max(an, fabs(a));
                }
        }
        std::vector<double> l0((size_t)B * nb), l1((size_t)B * nb);
        CK(cudaMemcpy(l0.data(), dld0, l0.size() * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(l1.data(), dld1, l1.size() * 8, cudaMemcpyDeviceToHost));
        double ldd = 0;
        for (int c = 0; c < B; c++) { if (variant == 2 && c % 3 == 0) continue; for (int k = 0; k < nb; k++) ldd = std::max(ldd, fabs(l0[c * nb + k] - l1[c * nb + k])); }
        double invd = 0, invref = 0;
        if (scaled || syrk) {
            std::vector<double> i0((size_t)B * nb * 4096), i1((size_t)B * nb * 4096);
            CK(cudaMemcpy(i0.data(), dinv0, i0.size() * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(i1.data(), dinv1, i1.size() * 8, cudaMemcpyDeviceToHost));
            for (int c = 0; c < B; c++) { if (variant == 2 && c % 3 == 0) continue; for (size_t e = 0; e < (size_t)nb * 4096; e++) { invd = std::max(invd, fabs(i0[c * nb * 4096 + e] - i1[c * nb * 4096 + e])); invref = std::max(invref, fabs(i0[c * nb * 4096 + e])); } }
        }
        std::vector<int> hst(B);
        CK(cudaMemcpy(hst.data(), dstatus, B * 4, cudaMemcpyDeviceToHost));
        int nfail = 0; for (int v : hst) nfail += v != 0;
        const double nact = variant == 2 ? B - (B + 2) / 3 : B;
        const double gflop = (syrk ? 2.0 : 1.0) * nact * (double)np * np * np / 3.0 / 1e9;
        printf("variant %d: new vs old max rel %.3e (abs %.3e)  upper %.1e  recon %.3e  logdet diff %.2e  inv diff %.2e (max %.2e)  failed %d\n",
               variant, rel, wabs, upper, rec / an, ldd, invd, invref, nfail);
        printf("           old %.3f ms (%.2f TFLOP/s)   new %.3f ms (%.2f TFLOP/s)   speed-up %.2fx\n", ms_old, gflop / ms_old, ms_new,
               gflop / ms_new, ms_old / ms_new);
    }
    return 0;
}
