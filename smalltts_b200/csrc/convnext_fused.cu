// Fused ConvNeXt-1D layer for the vocoder's HBM-bound tail (C = 32, 64; hf:263-297).
//
//   y   = x + gamma * (dwconv7_causal(rmsnorm(x)) + conv_b)
//   out = y + ffn_gamma * (W2 gelu(W1 rmsnorm'(y) + b1) + b2)
//
// HBM traffic per layer is exactly the algorithmic minimum: x is read once (fp32) and out is written once (fp32);
// the normalised bf16 operand, the 4C-wide hidden activation and y never leave the SM.  One persistent CTA per SM
// walks 128-row tiles; five warp roles overlap consecutive tiles through mbarrier pipelines:
//
//   warps 0-3   mixer : cp.async prefetch of the next x tile (+6 causal halo rows, zero-filled), RMSNorm, depthwise
//                       conv in place (fp32 y stays in shared memory), second RMSNorm -> bf16 A operand written in
//                       the UMMA K-major 128B-swizzled layout
//   warp  4     MMA   : tcgen05.mma  H[128 x 4C] = A W1^T  and, per 64-wide hidden chunk, O[128 x C] += G W2^T;
//                       accumulators in TMEM (two 256-column buffers; O aliases the first C columns of H, which the
//                       GELU warps have consumed by then); W1/W2 stay resident in shared memory (TMA-loaded once)
//   warps 5-12  GELU  : tcgen05.ld H chunk -> +b1 -> GELU -> bf16 -> swizzled shared-memory G chunk (double buffer)
//   warps 13-16 out   : tcgen05.ld O -> y + ffn_gamma*(O + b2) -> shared -> coalesced fp32 (and optional bf16) store
#include <cuda.h>

#include <mutex>

#include "kernels.cuh"
#include "ptx.cuh"

namespace stts {

namespace {

constexpr int TM = 128;  // rows per tile
constexpr int HALO = 6;  // causal context of the k=7 depthwise conv
constexpr int XR = TM + HALO;
constexpr int kThreadsFused = 17 * 32;

struct FusedParams {
  const float* x;
  float* out;
  bf16* out_bf16;  // optional bf16 copy (feeds the next transposed conv)
  int B, T;
  const float *norm_w, *conv_w, *conv_b, *gamma, *ffn_norm_w, *b1, *b2, *ffn_gamma;
  float eps;
};

template <int C>
struct FC {
  static constexpr int HID = 4 * C;
  static constexpr int NCH = HID / 64;  // 64-wide hidden chunks
  static constexpr int XP = C + 4;      // fp32 tile pitch: 16-byte aligned rows, conflict-free float4 row access
  static constexpr int W1_BYTES = HID * 128;      // [HID rows][64 k] bf16, K zero-padded to 64
  static constexpr int W2_BYTES = NCH * C * 128;  // NCH chunks of [C rows][64 k]
  static constexpr int A_BYTES = TM * 128;
  static constexpr int G_BYTES = TM * 128;
  static constexpr int X_BYTES = ((XR * XP * 4 + 127) / 128) * 128;
  // vectors: b1[HID] b2[C] ffn_gamma[C] norm_w[C] ffn_norm_w[C] gamma[C] conv_b[C] conv_w[7][C]
  static constexpr int VEC_FLOATS = HID + 6 * C + 7 * C;
  static constexpr int OFF_W1 = 0;
  static constexpr int OFF_W2 = OFF_W1 + W1_BYTES;
  static constexpr int OFF_A = OFF_W2 + W2_BYTES;
  static constexpr int OFF_G = OFF_A + 2 * A_BYTES;
  static constexpr int OFF_X = OFF_G + 2 * G_BYTES;
  static constexpr int OFF_VEC = OFF_X + 2 * X_BYTES;
  static constexpr int OFF_INV = OFF_VEC + VEC_FLOATS * 4;
  static constexpr int OFF_BAR = ((OFF_INV + XR * 4 + 15) / 16) * 16;
  static constexpr int SMEM = OFF_BAR + 17 * 8 + 16 + 1024;
};

__device__ __forceinline__ float gelu_fast(float x) {  // same fit as gemm.cu: max abs err 2.6e-5 vs erf GELU
  const float xc = fminf(fmaxf(x, -8.0f), 8.0f);
  const float x2 = xc * xc;
  const float p = fmaf(x2, fmaf(x2, 1.0153833e-3f, -0.10678167f), -2.3011139f);
  return __fdividef(x, 1.0f + exp2f(xc * p));
}
__device__ __forceinline__ uint32_t bf2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
// byte offset of element (row r, column k) in a [rows][64] bf16 K-major tile with 128-byte swizzle
__device__ __forceinline__ uint32_t sw128_off(int r, int k) {
  return static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 3) ^ (r & 7))) << 4) + (k & 7) * 2);
}

template <int C>
__global__ void __launch_bounds__(kThreadsFused, 1)
convnext_fused_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                      const FusedParams p) {
  using F = FC<C>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* W1s = smem + F::OFF_W1;
  uint8_t* W2s = smem + F::OFF_W2;
  float* vec = reinterpret_cast<float*>(smem + F::OFF_VEC);
  float* b1s = vec;
  float* b2s = b1s + F::HID;
  float* gfs = b2s + C;
  float* nws = gfs + C;
  float* fws = nws + C;
  float* gms = fws + C;
  float* cbs = gms + C;
  float* cws = cbs + C;  // [7][C]
  float* inv1 = reinterpret_cast<float*>(smem + F::OFF_INV);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F::OFF_BAR);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;
  uint64_t* a_empty = bars + 3;
  uint64_t* h_full = bars + 5;
  uint64_t* g_full = bars + 7;
  uint64_t* g_empty = bars + 9;
  uint64_t* o_full = bars + 11;
  uint64_t* tm_empty = bars + 13;
  uint64_t* x_empty = bars + 15;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
  auto Abuf = [&](int i) { return smem + F::OFF_A + i * F::A_BYTES; };
  auto Gbuf = [&](int i) { return smem + F::OFF_G + i * F::G_BYTES; };
  auto Xbuf = [&](int i) { return reinterpret_cast<float*>(smem + F::OFF_X + i * F::X_BYTES); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_b = (p.T + TM - 1) / TM;
  const int ntiles = p.B * tiles_per_b;
  const int first = blockIdx.x, stride = gridDim.x;
  const int n_my = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;

  // ---------------- one-time setup
  for (int i = threadIdx.x; i < F::HID; i += kThreadsFused) b1s[i] = p.b1[i];
  for (int i = threadIdx.x; i < C; i += kThreadsFused) {
    b2s[i] = p.b2[i]; gfs[i] = p.ffn_gamma[i]; nws[i] = p.norm_w[i]; fws[i] = p.ffn_norm_w[i];
    gms[i] = p.gamma[i]; cbs[i] = p.conv_b[i];
  }
  for (int i = threadIdx.x; i < 7 * C; i += kThreadsFused) cws[i] = p.conv_w[(i % C) * 7 + i / C];  // tap-major
  for (int i = threadIdx.x; i < 2 * F::A_BYTES / 16; i += kThreadsFused) {
    reinterpret_cast<uint4*>(smem + F::OFF_A)[i] = make_uint4(0, 0, 0, 0);  // K padding (C = 32) stays zero
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 17; ++i) ptx::mbar_init(&bars[i], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 4) ptx::tmem_alloc<512>(tmem_slot);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ================================================================== mixer
    const int tid = threadIdx.x;
    constexpr int CV = C / 4;      // float4 per row
    constexpr int LPR = CV;        // lanes per row (16 or 8)
    constexpr int RPI = 32 / LPR;  // rows per warp iteration
    constexpr int NSEG = 128 / C;  // time segments per channel
    constexpr int SEGLEN = TM / NSEG;
    auto issue_load = [&](int tile, int buf) {
      const int b = tile / tiles_per_b, t0 = (tile % tiles_per_b) * TM;
      const float* src_b = p.x + static_cast<long long>(b) * p.T * C;
      const uint32_t dst0 = ptx::smem_u32(Xbuf(buf));
      for (int i = tid; i < XR * CV; i += 128) {
        const int r = i / CV, c4 = i % CV;
        const int t = t0 - HALO + r;
        const bool ok = (t >= 0) && (t < p.T);
        const float* src = src_b + static_cast<long long>(ok ? t : 0) * C + c4 * 4;
        ptx::cp_async_16(dst0 + (r * F::XP + c4 * 4) * 4, src, ok ? 16u : 0u);
      }
      ptx::cp_async_commit();
    };
    if (n_my > 0) issue_load(first, 0);
    for (int it = 0; it < n_my; ++it) {
      const int buf = it & 1;
      if (it + 1 < n_my) {
        // the other x buffer was last used by iteration it-1: wait until its out warps released it
        if (it >= 1) ptx::mbar_wait(&x_empty[buf ^ 1], ((it - 1) >> 1) & 1);
        issue_load(first + (it + 1) * stride, buf ^ 1);
        ptx::cp_async_wait<1>();
      } else {
        ptx::cp_async_wait<0>();
      }
      ptx::named_bar_sync(1, 128);
      float* xs = Xbuf(buf);
      // (1) 1/rms of every staged row
      for (int r0 = warp * RPI; r0 < XR; r0 += 4 * RPI) {
        const int r = r0 + lane / LPR, sl = lane % LPR;
        float s = 0.f;
        if (r < XR) {
          const float4 v = *reinterpret_cast<const float4*>(xs + r * F::XP + sl * 4);
          s = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
#pragma unroll
        for (int o = LPR >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (r < XR && sl == 0) inv1[r] = 1.0f / sqrtf(s / C + p.eps);
      }
      ptx::named_bar_sync(1, 128);
      // (2) depthwise causal conv along time, one (channel, segment) per thread, y written in place
      {
        const int c = tid % C, seg = tid / C;
        const int rs = HALO + seg * SEGLEN;
        float w[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) w[j] = cws[j * C + c];
        const float cb = cbs[c], gm = gms[c], nw = nws[c];
        float win[7];
        win[0] = 0.f;
#pragma unroll
        for (int j = 1; j < 7; ++j) {
          const int r = rs - 7 + j;
          win[j] = xs[r * F::XP + c] * inv1[r] * nw;
        }
        ptx::named_bar_sync(1, 128);  // every warm-up read precedes the in-place writes of the previous segment
#pragma unroll 4
        for (int r = rs; r < rs + SEGLEN; ++r) {
          const float xv = xs[r * F::XP + c];
#pragma unroll
          for (int j = 0; j < 6; ++j) win[j] = win[j + 1];
          win[6] = xv * inv1[r] * nw;
          float acc = cb;
#pragma unroll
          for (int j = 0; j < 7; ++j) acc = fmaf(w[j], win[j], acc);
          xs[r * F::XP + c] = fmaf(gm, acc, xv);
        }
      }
      ptx::named_bar_sync(1, 128);
      // (3) second RMSNorm -> bf16 A operand (UMMA K-major, 128B swizzle); A buffer must be free (MMA1 of it-2 done)
      if (it >= 2) ptx::mbar_wait(&a_empty[buf], ((it - 2) >> 1) & 1);
      uint8_t* As = Abuf(buf);
      for (int r0 = warp * RPI; r0 < TM; r0 += 4 * RPI) {
        const int r = r0 + lane / LPR, sl = lane % LPR;
        const float4 v = *reinterpret_cast<const float4*>(xs + (r + HALO) * F::XP + sl * 4);
        float s = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
#pragma unroll
        for (int o = LPR >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float inv = 1.0f / sqrtf(s / C + p.eps);
        const float4 fw = *reinterpret_cast<const float4*>(fws + sl * 4);
        uint2 pk;
        pk.x = bf2(v.x * inv * fw.x, v.y * inv * fw.y);
        pk.y = bf2(v.z * inv * fw.z, v.w * inv * fw.w);
        *reinterpret_cast<uint2*>(As + sw128_off(r, sl * 4)) = pk;
      }
      ptx::fence_proxy_async();
      ptx::named_bar_sync(1, 128);
      if (tid == 0) ptx::mbar_arrive(&a_full[buf]);
    }
  } else if (warp == 4) {
    // ================================================================== MMA issuer
    if (lane == 0 && n_my > 0) {
      ptx::mbar_expect_tx(w_full, F::W1_BYTES + F::W2_BYTES);
      ptx::tma_load_2d(W1s, &tmW1, w_full, 0, 0);
      for (int c = 0; c < F::NCH; ++c) ptx::tma_load_2d(W2s + c * C * 128, &tmW2, w_full, c * 64, 0);
      ptx::mbar_wait(w_full, 0);
      constexpr uint32_t idesc1 = ptx::umma_idesc_bf16(TM, F::HID);
      constexpr uint32_t idesc2 = ptx::umma_idesc_bf16(TM, C);
      const uint64_t dW1 = ptx::umma_desc_sw128(ptx::smem_u32(W1s));
      auto mma1 = [&](int it) {
        const int buf = it & 1;
        ptx::mbar_wait(&a_full[buf], (it >> 1) & 1);
        if (it >= 2) ptx::mbar_wait(&tm_empty[buf], ((it - 2) >> 1) & 1);  // out warps drained this TMEM buffer
        ptx::tc_fence_after();
        const uint64_t dA = ptx::umma_desc_sw128(ptx::smem_u32(Abuf(buf)));
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem_base + buf * 256, dA + 2 * k, dW1 + 2 * k, idesc1, k > 0);
        ptx::umma_commit(&a_empty[buf]);
        ptx::umma_commit(&h_full[buf]);
      };
      mma1(0);
      uint32_t gcount = 0;
      for (int it = 0; it < n_my; ++it) {
        if (it + 1 < n_my) mma1(it + 1);  // run one tile ahead so that the GELU warps always have work
        for (int c = 0; c < F::NCH; ++c, ++gcount) {
          const int gb = gcount & 1;
          ptx::mbar_wait(&g_full[gb], (gcount >> 1) & 1);
          ptx::tc_fence_after();
          const uint64_t dG = ptx::umma_desc_sw128(ptx::smem_u32(Gbuf(gb)));
          const uint64_t dW2 = ptx::umma_desc_sw128(ptx::smem_u32(W2s + c * C * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            ptx::umma_bf16(tmem_base + (it & 1) * 256, dG + 2 * k, dW2 + 2 * k, idesc2, (c | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(&g_empty[gb]);
        }
        ptx::umma_commit(&o_full[it & 1]);
      }
    }
  } else if (warp < 13) {
    // ================================================================== GELU: H (TMEM) -> G (smem, bf16)
    const int gw = warp - 5;
    const int q = warp & 3;      // TMEM lane quarter this warp may access
    const int half = gw >> 2;    // which 32 of the chunk's 64 columns
    const int r = q * 32 + lane;
    uint32_t gcount = 0;
    for (int it = 0; it < n_my; ++it) {
      ptx::mbar_wait(&h_full[it & 1], (it >> 1) & 1);
      ptx::tc_fence_after();
      for (int c = 0; c < F::NCH; ++c, ++gcount) {
        const int gb = gcount & 1;
        if (gcount >= 2) ptx::mbar_wait(&g_empty[gb], ((gcount - 2) >> 1) & 1);
        uint32_t rr[32];
        ptx::tmem_ld_32x32(tmem_base + (it & 1) * 256 + (static_cast<uint32_t>(q * 32) << 16) + c * 64 + half * 32, rr);
        ptx::tmem_ld_wait();
        const float* bb = b1s + c * 64 + half * 32;
        uint8_t* Gs = Gbuf(gb);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 ba = *reinterpret_cast<const float4*>(bb + 8 * j);
          const float4 bc = *reinterpret_cast<const float4*>(bb + 8 * j + 4);
          uint4 pk;
          pk.x = bf2(gelu_fast(__uint_as_float(rr[8 * j + 0]) + ba.x), gelu_fast(__uint_as_float(rr[8 * j + 1]) + ba.y));
          pk.y = bf2(gelu_fast(__uint_as_float(rr[8 * j + 2]) + ba.z), gelu_fast(__uint_as_float(rr[8 * j + 3]) + ba.w));
          pk.z = bf2(gelu_fast(__uint_as_float(rr[8 * j + 4]) + bc.x), gelu_fast(__uint_as_float(rr[8 * j + 5]) + bc.y));
          pk.w = bf2(gelu_fast(__uint_as_float(rr[8 * j + 6]) + bc.z), gelu_fast(__uint_as_float(rr[8 * j + 7]) + bc.w));
          *reinterpret_cast<uint4*>(Gs + sw128_off(r, (half * 4 + j) * 8)) = pk;
        }
        ptx::fence_proxy_async();
        ptx::tc_fence_before();
        ptx::named_bar_sync(2, 256);
        if (gw == 0 && lane == 0) ptx::mbar_arrive(&g_full[gb]);
      }
    }
  } else {
    // ================================================================== out: O (TMEM) + y -> global
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int otid = threadIdx.x - 13 * 32;
    constexpr int CV = C / 4;
    for (int it = 0; it < n_my; ++it) {
      const int buf = it & 1;
      const int tile = first + it * stride;
      ptx::mbar_wait(&o_full[buf], (it >> 1) & 1);
      ptx::tc_fence_after();
      float* xs = Xbuf(buf);
      float* yrow = xs + (r + HALO) * F::XP;
#pragma unroll
      for (int cc = 0; cc < C / 32; ++cc) {
        uint32_t rr[32];
        ptx::tmem_ld_32x32(tmem_base + buf * 256 + (static_cast<uint32_t>(q * 32) << 16) + cc * 32, rr);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int col = cc * 32 + 4 * j;
          float4 y = *reinterpret_cast<const float4*>(yrow + col);
          const float4 b2 = *reinterpret_cast<const float4*>(b2s + col);
          const float4 gf = *reinterpret_cast<const float4*>(gfs + col);
          y.x = fmaf(gf.x, __uint_as_float(rr[4 * j + 0]) + b2.x, y.x);
          y.y = fmaf(gf.y, __uint_as_float(rr[4 * j + 1]) + b2.y, y.y);
          y.z = fmaf(gf.z, __uint_as_float(rr[4 * j + 2]) + b2.z, y.z);
          y.w = fmaf(gf.w, __uint_as_float(rr[4 * j + 3]) + b2.w, y.w);
          *reinterpret_cast<float4*>(yrow + col) = y;
        }
      }
      ptx::tc_fence_before();
      ptx::named_bar_sync(3, 128);
      if (otid == 0) ptx::mbar_arrive(&tm_empty[buf]);  // TMEM buffer may be overwritten by MMA1 of tile it+2
      // coalesced store of the finished rows
      const int b = tile / tiles_per_b, t0 = (tile % tiles_per_b) * TM;
      const int nrows = min(TM, p.T - t0);
      const long long obase = (static_cast<long long>(b) * p.T + t0) * C;
      float4* dst = reinterpret_cast<float4*>(p.out + obase);
      uint2* dstb = p.out_bf16 ? reinterpret_cast<uint2*>(p.out_bf16 + obase) : nullptr;
      for (int i = otid; i < nrows * CV; i += 128) {
        const int rr2 = i / CV, c4 = i % CV;
        const float4 v = *reinterpret_cast<const float4*>(xs + (rr2 + HALO) * F::XP + c4 * 4);
        dst[i] = v;
        if (dstb) {
          uint2 pk;
          pk.x = bf2(v.x, v.y);
          pk.y = bf2(v.z, v.w);
          dstb[i] = pk;
        }
      }
      ptx::named_bar_sync(3, 128);
      if (otid == 0) ptx::mbar_arrive(&x_empty[buf]);  // x buffer may be refilled for tile it+2
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) ptx::tmem_dealloc<512>(tmem_base);
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeFn>(p);
    }
  });
  return fn;
}

bool tmap_2d(CUtensorMap* out, const void* ptr, uint64_t cols, uint64_t rows, uint32_t box_rows) {
  EncodeFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t str[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t es[2] = {1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, str, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int C>
cudaError_t launch_fused(cudaStream_t st, const FusedParams& p, const bf16* w1, const bf16* w2, int num_sms) {
  static bool once = false;
  if (!once) {
    cudaError_t e = cudaFuncSetAttribute(convnext_fused_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, FC<C>::SMEM);
    if (e != cudaSuccess) return e;
    once = true;
  }
  CUtensorMap m1, m2;
  if (!tmap_2d(&m1, w1, C, 4 * C, 4 * C)) return cudaErrorInvalidValue;  // W1 [4C, C]: one box of all rows, K padded to 64
  if (!tmap_2d(&m2, w2, 4 * C, C, C)) return cudaErrorInvalidValue;      // W2 [C, 4C]: one box per 64-wide K chunk
  const int ntiles = p.B * ((p.T + TM - 1) / TM);
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  convnext_fused_kernel<C><<<grid, kThreadsFused, FC<C>::SMEM, st>>>(m1, m2, p);
  ++g_launch_count;
  return cudaGetLastError();
}

}  // namespace

cudaError_t convnext_fused(cudaStream_t st, const float* x, int B, int T, int C, const float* norm_w,
                           const float* conv_w, const float* conv_b, const float* gamma, const float* ffn_norm_w,
                           const bf16* w1, const float* b1, const bf16* w2, const float* b2, const float* ffn_gamma,
                           float eps, float* out, bf16* out_bf16) {
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  FusedParams p;
  p.x = x; p.out = out; p.out_bf16 = out_bf16; p.B = B; p.T = T;
  p.norm_w = norm_w; p.conv_w = conv_w; p.conv_b = conv_b; p.gamma = gamma; p.ffn_norm_w = ffn_norm_w;
  p.b1 = b1; p.b2 = b2; p.ffn_gamma = ffn_gamma; p.eps = eps;
  if (C == 64) return launch_fused<64>(st, p, w1, w2, num_sms);
  if (C == 32) return launch_fused<32>(st, p, w1, w2, num_sms);
  return cudaErrorInvalidValue;
}

}  // namespace stts
