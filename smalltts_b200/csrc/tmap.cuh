// Host-side CUtensorMap encoding without linking libcuda: cuTensorMapEncodeTiled is resolved through the runtime's
// driver entry point query, so the library cross-compiles and links on a machine without a GPU driver.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

namespace stts {

using TmapEncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline TmapEncodeFn tmap_encode_fn() {
  static TmapEncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<TmapEncodeFn>(p);
    }
  });
  return fn;
}

// rank <= 4 tiled map, element strides 1, out-of-bounds elements read as zero.
// dims[0] is the contiguous dimension; strides_bytes[i] is the pitch of dims[i + 1].
inline bool tmap_tiled(CUtensorMap* out, CUtensorMapDataType dtype, const void* ptr, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  TmapEncodeFn fn = tmap_encode_fn();
  if (fn == nullptr || rank < 1 || rank > 4) return false;
  cuuint64_t gdim[4];
  cuuint64_t gstr[3];
  cuuint32_t bx[4];
  cuuint32_t es[4] = {1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  return fn(out, dtype, rank, const_cast<void*>(ptr), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace stts
