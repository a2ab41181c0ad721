// Host side of TMA: tensor-map (CUtensorMap) construction for the row-major fp32 matrices the kernels stream.
// cuTensorMapEncodeTiled is resolved through the runtime (no link-time dependency on libcuda).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vmtl {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled tma_encode_fn() {
  static PFN_encodeTiled fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<PFN_encodeTiled>(p);
  }();
  return fn;
}

// Row-major fp32 [rows, cols] matrix, box = [box_rows x 32 floats] (one 128-byte swizzle atom wide), rows
// outside the matrix read as 0 / are clipped on store.  mn32 = false: SWIZZLE_128B (K-major operands, staging
// tiles); mn32 = true: SWIZZLE_128B_ATOM_32B, the layout MN-major tf32 operands require.
inline bool make_tmap_2d_sw(CUtensorMap* m, const float* base, int64_t rows, int cols, int box_rows, bool mn32) {
  PFN_encodeTiled enc = tma_encode_fn();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
  const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUtensorMapSwizzle sw = mn32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline bool make_tmap_2d(CUtensorMap* m, const float* base, int64_t rows, int cols, int box_rows) {
  return make_tmap_2d_sw(m, base, rows, cols, box_rows, false);
}

}  // namespace vmtl
