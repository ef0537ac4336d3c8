// TMA tensor maps (cuTensorMapEncodeTiled) for the lattice fields of the strip kernel.  The driver entry point is
// fetched through the runtime (cudaGetDriverEntryPoint), so the library does not link libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>

namespace s2d {

typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline tmap_encode_fn tmap_encoder() {
  static tmap_encode_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p ||
        q != cudaDriverEntryPointSuccess)
      throw std::runtime_error("cuTensorMapEncodeTiled is not available from this driver");
    fn = (tmap_encode_fn)p;
  }
  return fn;
}

// rank-3 (or rank-2 with nz == 1 plane count 0) map of a lattice field: dims (LXP, LZ[, ncomp]), row pitch LXP
// elements, component stride `comp_stride` elements; box (bw, brows[, ncomp]); no swizzle, zero fill out of bounds
inline CUtensorMap lattice_tmap(void* base, int elem_bytes, size_t LXP, size_t LZ, int ncomp, size_t comp_stride, int bw,
                                int brows) {
  CUtensorMap m;
  const int rank = ncomp > 0 ? 3 : 2;
  cuuint64_t dims[3] = {(cuuint64_t)LXP, (cuuint64_t)LZ, (cuuint64_t)(ncomp > 0 ? ncomp : 1)};
  cuuint64_t strides[2] = {(cuuint64_t)LXP * elem_bytes, (cuuint64_t)comp_stride * elem_bytes};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)brows, (cuuint32_t)(ncomp > 0 ? ncomp : 1)};
  cuuint32_t es[3] = {1, 1, 1};
  const CUtensorMapDataType dt = elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const CUresult rc = tmap_encoder()(&m, dt, (cuuint32_t)rank, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)rc));
  return m;
}

}  // namespace s2d
