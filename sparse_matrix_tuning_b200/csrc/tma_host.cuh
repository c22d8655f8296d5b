// Host-side construction of the TMA descriptors all tensor-core kernels of this library use: a 2-D map over a row-major
// 16-bit matrix, box = {64 elements of the contiguous dimension (128 bytes), box_rows rows}, 128-byte swizzle, zero fill
// outside the matrix.  `cuTensorMapEncodeTiled` is resolved through the runtime (no link dependency on libcuda).
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace smt {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn tma_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// `cols` = extent of the contiguous dimension, `rows` = extent of the strided one, `ld` = row pitch in elements.
inline int encode_2d_sw128(CUtensorMap* map, const void* base, int64_t cols, int64_t rows, int64_t ld, int dtype,
                           int box_rows, const char* who) {
  EncodeTiledFn enc = tma_encode_fn();
  if (!enc) {
    set_error("%s: cuTensorMapEncodeTiled not available from the driver", who);
    return SMT_ERR_CUDA;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dtype == SMT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                   const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled failed with CUresult %d", who, (int)r);
    return SMT_ERR_CUDA;
  }
  return SMT_OK;
}

}  // namespace smt
