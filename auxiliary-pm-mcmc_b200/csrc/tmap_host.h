// tmap_host.h -- host side of the TMA path: tensor maps over the batched fp64 matrix buffers.
// cuTensorMapEncodeTiled is fetched through the runtime (cudaGetDriverEntryPoint), so nothing links against libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace apm {

typedef CUresult (*PFN_apm_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_apm_encodeTiled tmap_encode_fn() {
    static PFN_apm_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_apm_encodeTiled)p;
    }
    return fn;
}

// 2-D view of `nmat` stacked row-major fp64 matrices [nmat][np][np]: dim0 = column, dim1 = stacked row (m * np + r).
// Box = 16 columns (one 128-byte swizzle row) x box_rows rows, SWIZZLE_128B: the operand stages of k_chol_flow (64 rows: row-major
// operands with k contiguous; 16 rows: k-major 16x16 boxes of L_K for the fused M' = I + L_K^T W L_K source).
inline bool make_matrix_tmap(CUtensorMap* out, const double* base, int np, long long nmat, int box_rows = 64) {
    PFN_apm_encodeTiled enc = tmap_encode_fn();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)np, (cuuint64_t)np * (cuuint64_t)nmat};
    const cuuint64_t strides[1] = {(cuuint64_t)np * sizeof(double)};
    const cuuint32_t box[2] = {16, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace apm
