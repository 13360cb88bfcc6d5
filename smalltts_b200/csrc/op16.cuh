// The 16-bit operand type of the tensor-core GEMMs and the attention kernel.
//
//   default build          : bf16 (8-bit significand, fp32's exponent range)            -> libsmalltts_b200.so
//   -DSTTS_OPERAND_F16     : fp16 (11-bit significand -- the precision of TF32 -- at half of TF32's bytes)
//                                                                                       -> libsmalltts_b200_tight.so
// Accumulation is fp32 in both builds, and so are the residual streams, norms, softmax, RoPE and the sampler; only the
// rounding of GEMM / attention operands changes.  The "tight" build is the parity mode (SURVEY 7 "Tolerance vs tensor
// cores"): same kernels, same memory traffic, about 8x smaller operand rounding error.  fp16's range (65504) is wide
// enough for this network: every operand is a normalised activation, a probability, a weight, or a latent of O(1..100).
// The type keeps the name `bf16` throughout the sources (it is what the default build uses).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace stts {

#ifdef STTS_OPERAND_F16
typedef __half bf16;
#define STTS_OP16_NAME "fp16"
#else
typedef __nv_bfloat16 bf16;
#define STTS_OP16_NAME "bf16"
#endif

// two floats -> packed pair of 16-bit operands (round to nearest even), as raw bits
__host__ __device__ __forceinline__ uint32_t op16_pack2(float a, float b) {
#ifdef STTS_OPERAND_F16
  __half2 p = __floats2half2_rn(a, b);
#else
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
#endif
  return *reinterpret_cast<uint32_t*>(&p);
}
__host__ __device__ __forceinline__ bf16 op16_from_float(float a) {
#ifdef STTS_OPERAND_F16
  return __float2half_rn(a);
#else
  return __float2bfloat16_rn(a);
#endif
}
__host__ __device__ __forceinline__ float op16_to_float(bf16 a) {
#ifdef STTS_OPERAND_F16
  return __half2float(a);
#else
  return __bfloat162float(a);
#endif
}
// packed pair (raw bits) -> two floats
__device__ __forceinline__ float2 op16_unpack2(uint32_t bits) {
#ifdef STTS_OPERAND_F16
  return __half22float2(*reinterpret_cast<const __half2*>(&bits));
#else
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&bits));
#endif
}

}  // namespace stts
