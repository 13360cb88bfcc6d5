// Fused ConvNeXt feed-forward for the vocoder's C = 128 stage (hf:293-297):
//
//   out = y + ffn_gamma * (W2 gelu(W1 a + b1) + b2)        a = bf16 rmsnorm'(y) from the token mixer, y fp32
//
// The unfused path writes the 4C-wide hidden activation to HBM and reads it back (2 x 491 MB per layer at B = 8 x
// 10 s, against 613 MB for everything else).  Here it never leaves the SM: per 128-row tile the hidden layer is
// produced 64 columns at a time in TMEM, passed through GELU into shared memory and immediately consumed by the second
// GEMM.  W1 / W2 (256 KB) do not fit in shared memory next to the tiles, so they stream from L2 through a 3-stage
// TMA ring, one 64-column hidden chunk (16 KB of W1 + 16 KB of W2) per stage.
//
//   warp  0     TMA producer : A tile (2 K-atoms, double buffered) + weight-chunk ring; the first weight stages are
//                              requested before the PDL dependency wait (weights never depend on the predecessor)
//   warp  1     MMA1 issuer  : H[c] = A W1c^T (8 x tcgen05.mma N=64), runs ahead as far as the two H buffers allow
//   warps 2-9   GELU         : tcgen05.ld H chunk -> +b1 -> packed-fp16 2*gelu -> swizzled G chunk (double buffered)
//   warps 10-13 out          : tcgen05.ld O -> transposition buffer -> y + ffn_gamma*(O + b2) on 128-byte row segments
//   warp  14    MMA2 issuer  : O += G[c] W2c^T (4 x N=128).  Two issuing warps because one thread's wait -> issue ->
//                              commit chain (~400 clk per step) would otherwise pace the whole tile.
#include <cuda.h>
#include <cuda_fp16.h>

#include "kernels.cuh"
#include "launch.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace stts {

namespace {

constexpr int TM = 128;   // rows per tile
constexpr int HC = 64;    // hidden columns per chunk
constexpr int kWStages = 3;
constexpr int kGeluWarp0 = 2, kOutWarp0 = 10;
constexpr int kMma2Warp = 14;
constexpr int kThreadsFfn = 15 * 32;

struct FfnParams {
  const float* y;     // residual rows [M, C]
  float* out;         // [M, C]
  bf16* out_bf16;     // optional bf16 copy of out
  int M;
  const float *b1, *b2, *ffn_gamma;
};

template <int C>
struct FF {
  static constexpr int HID = 4 * C;
  static constexpr int NCH = HID / HC;           // hidden chunks per tile
  static constexpr int KA = C / 64;              // 64-wide K atoms of the A operand
  static constexpr int A_BYTES = KA * TM * 128;  // one A tile
  static constexpr int W1C_BYTES = KA * HC * 128;
  static constexpr int W2C_BYTES = C * 128;
  static constexpr int W_BYTES = W1C_BYTES + W2C_BYTES;
  static constexpr int G_BYTES = TM * 128;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_W = OFF_A + 2 * A_BYTES;
  static constexpr int OFF_G = OFF_W + kWStages * W_BYTES;
  static constexpr int OFF_STG = OFF_G + 2 * G_BYTES;  // 4 x 4 KB transposition staging (out warps)
  static constexpr int OFF_VEC = OFF_STG + 4 * 4096;    // b1 as fp16 [HID] | ffn_gamma*b2 [C] | ffn_gamma [C]
  static constexpr int OFF_BAR = OFF_VEC + HID * 2 + 2 * C * 4;
  static constexpr int NBAR = 2 + 2 + kWStages + kWStages + 2 + 2 + 2 + 2 + 2 + 2;
  static constexpr int SMEM = OFF_BAR + NBAR * 8 + 16 + 1024;
  static constexpr uint32_t TM_H = 0;           // TMEM columns: H[2] x HC, O[2] x C
  static constexpr uint32_t TM_O = 2 * HC;
  static_assert(TM_O + 2 * C <= 512, "accumulators do not fit in TMEM");
  static_assert(SMEM <= 232448, "fused FFN tile does not fit in shared memory");
};

// x (1 + tanh(x (A + B x^2))) on two values in packed fp16, see gemm.cu / convnext_fused.cu
__device__ __forceinline__ uint32_t gelu2_half2(float a, float b, __half2 bias) {
  const __half2 x = __hadd2(__floats2half2_rn(a, b), bias);
  const __half2 p = __hfma2(__hmul2(x, x), __float2half2_rn(0.03470089f), __float2half2_rn(0.80015708f));
  const __half2 u = __hmul2(x, p);
  uint32_t ui = *reinterpret_cast<const uint32_t*>(&u), ti;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(ti) : "r"(ui));
  const __half2 t = *reinterpret_cast<const __half2*>(&ti);
  const __half2 g = __hfma2(x, t, x);
  return *reinterpret_cast<const uint32_t*>(&g);
}
__device__ __forceinline__ uint32_t bf2(float a, float b) {
  return op16_pack2(a, b);
}
// byte offset of element (row r, column k) in a [rows][64] 16-bit K-major tile with 128-byte swizzle
__device__ __forceinline__ uint32_t sw128_off(int r, int k) {
  return static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 3) ^ (r & 7))) << 4) + (k & 7) * 2);
}

template <int C>
__global__ void __launch_bounds__(kThreadsFfn, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const FfnParams p) {
  using F = FF<C>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  __half* b1h = reinterpret_cast<__half*>(smem + F::OFF_VEC);
  float* b2g = reinterpret_cast<float*>(smem + F::OFF_VEC + F::HID * 2);
  float* fgs = b2g + C;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F::OFF_BAR);
  uint64_t* a_full = bars;               // [2] TMA -> MMA
  uint64_t* a_empty = bars + 2;          // [2] MMA -> TMA (all MMA1 of the tile issued and complete)
  uint64_t* w_full = bars + 4;           // [kWStages]
  uint64_t* w_empty = w_full + kWStages; // [kWStages] MMA2 of the chunk complete
  uint64_t* h_full = w_empty + kWStages; // [2] MMA1 chunk complete
  uint64_t* h_empty = h_full + 2;        // [2] 8 GELU warps have read the chunk out of TMEM
  uint64_t* g_full = h_empty + 2;        // [2] 8 GELU warps have written the G chunk
  uint64_t* g_empty = g_full + 2;        // [2] MMA2 has read the G chunk
  uint64_t* o_full = g_empty + 2;        // [2] last MMA2 of the tile complete
  uint64_t* o_empty = o_full + 2;        // [2] 4 out warps have read O out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);
  auto Abuf = [&](int i) { return smem + F::OFF_A + i * F::A_BYTES; };
  auto Wbuf = [&](int i) { return smem + F::OFF_W + i * F::W_BYTES; };
  auto Gbuf = [&](int i) { return smem + F::OFF_G + i * F::G_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = (p.M + TM - 1) / TM;
  const int first = blockIdx.x, stride = gridDim.x;
  const int n_my = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;

  // ---------------- one-time setup (weights only: may overlap the predecessor's tail under PDL)
  for (int i = threadIdx.x; i < F::HID; i += kThreadsFfn) b1h[i] = __float2half_rn(p.b1[i]);
  for (int i = threadIdx.x; i < C; i += kThreadsFfn) {
    b2g[i] = p.ffn_gamma[i] * p.b2[i];
    fgs[i] = p.ffn_gamma[i];
  }
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW1);
    ptx::prefetch_tmap(&tmW2);
    for (int i = 0; i < F::NBAR; ++i) {
      const bool eight = (&bars[i] >= h_empty && &bars[i] < g_empty);  // h_empty[2], g_full[2]: one arrive per GELU warp
      const bool four = (&bars[i] >= o_empty);                           // o_empty[2]: one arrive per out warp
      ptx::mbar_init(&bars[i], eight ? 8u : (four ? 4u : 1u));
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_trigger();

  if (warp == 0) {
    // ================================================================== TMA producer
    auto issue_w = [&](int c, int ws) {  // hidden chunk c of the weights into ring stage ws (one elected lane)
      ptx::mbar_expect_tx(&w_full[ws], F::W_BYTES);
#pragma unroll
      for (int ka = 0; ka < F::KA; ++ka) ptx::tma_load_2d(Wbuf(ws) + ka * (HC * 128), &tmW1, &w_full[ws], ka * 64, c * HC);
      ptx::tma_load_2d(Wbuf(ws) + F::W1C_BYTES, &tmW2, &w_full[ws], c * HC, 0);
    };
    const int total_chunks = n_my * F::NCH;
    const int w_pre = total_chunks < kWStages ? total_chunks : kWStages;  // requested ahead of the dependency wait
    if (ptx::elect_one()) {
      for (int i = 0; i < w_pre; ++i) issue_w(i % F::NCH, i);
    }
    __syncwarp();
    ptx::pdl_wait();
    uint32_t ws = 0, wph = 0;
    int gc = 0;
    for (int tl = 0; tl < n_my; ++tl) {
      const int tile = first + tl * stride;
      const int ab = tl & 1;
      ptx::mbar_wait(&a_empty[ab], ((tl >> 1) & 1) ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(&a_full[ab], F::A_BYTES);
#pragma unroll
        for (int ka = 0; ka < F::KA; ++ka) ptx::tma_load_2d(Abuf(ab) + ka * (TM * 128), &tmA, &a_full[ab], ka * 64, tile * TM);
      }
      __syncwarp();
      for (int c = 0; c < F::NCH; ++c, ++gc) {
        if (gc >= w_pre) {
          ptx::mbar_wait(&w_empty[ws], wph ^ 1);
          if (ptx::elect_one()) issue_w(c, ws);
          __syncwarp();
        }
        if (++ws == kWStages) { ws = 0; wph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA1 issuer: H chunks
    constexpr uint32_t idesc1 = ptx::umma_idesc_bf16(TM, HC);
    uint32_t ws = 0, wph = 0;
    int gc = 0;  // global chunk counter (H buffer = gc & 1, use count = gc >> 1)
    for (int tl = 0; tl < n_my; ++tl) {
      const int ab = tl & 1;
      ptx::mbar_wait(&a_full[ab], (tl >> 1) & 1);
      const uint64_t dA = ptx::umma_desc_sw128(ptx::smem_u32(Abuf(ab)));
      for (int c = 0; c < F::NCH; ++c, ++gc) {
        const int hb = gc & 1, n = gc >> 1;
        ptx::mbar_wait(&w_full[ws], wph);
        ptx::mbar_wait(&h_empty[hb], (n & 1) ^ 1);  // GELU warps have drained this H buffer
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint64_t dW = ptx::umma_desc_sw128(ptx::smem_u32(Wbuf(ws)));
#pragma unroll
          for (int ka = 0; ka < F::KA; ++ka) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              ptx::umma_bf16(tmem_base + F::TM_H + hb * HC, dA + ka * ((TM * 128) >> 4) + 2 * k,
                             dW + ka * ((HC * 128) >> 4) + 2 * k, idesc1, (ka | k) != 0 ? 1u : 0u);
            }
          }
          ptx::umma_commit(&h_full[hb]);
          if (c + 1 == F::NCH) ptx::umma_commit(&a_empty[ab]);  // every MMA1 of the tile issued: A is free when they complete
        }
        __syncwarp();
        if (++ws == kWStages) { ws = 0; wph ^= 1; }
      }
    }
  } else if (warp == kMma2Warp) {
    // ================================================================== MMA2 issuer: O += G W2c^T
    // MMA2 operands (G and W2) are fp16: clear the two bf16 format fields of the descriptor
    constexpr uint32_t idesc2 = ptx::umma_idesc_bf16(TM, C) & ~((1u << 7) | (1u << 10));
    uint32_t ws = 0, wph = 0;
    int gc = 0;
    for (int tl = 0; tl < n_my; ++tl) {
      const int ob = tl & 1;
      ptx::mbar_wait(&o_empty[ob], ((tl >> 1) & 1) ^ 1);  // out warps have drained O of tile tl-2
      for (int c = 0; c < F::NCH; ++c, ++gc) {
        const int gb = gc & 1, n = gc >> 1;
        ptx::mbar_wait(&w_full[ws], wph);  // already complete (MMA1 of this chunk waited for it); orders the W2c read
        ptx::mbar_wait(&g_full[gb], n & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint64_t dG = ptx::umma_desc_sw128(ptx::smem_u32(Gbuf(gb)));
          const uint64_t dW2 = ptx::umma_desc_sw128(ptx::smem_u32(Wbuf(ws) + F::W1C_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            ptx::umma_bf16(tmem_base + F::TM_O + ob * C, dG + 2 * k, dW2 + 2 * k, idesc2, (c | k) != 0 ? 1u : 0u);
          }
          // MMA1 of this chunk completed long ago (its result went through the GELU warps), so this commit alone
          // releases the weight stage
          ptx::umma_commit(&w_empty[ws]);
          ptx::umma_commit(&g_empty[gb]);
          if (c + 1 == F::NCH) ptx::umma_commit(&o_full[ob]);
        }
        __syncwarp();
        if (++ws == kWStages) { ws = 0; wph ^= 1; }
      }
    }
  } else if (warp < kOutWarp0) {
    // ================================================================== GELU: H (TMEM) -> G (smem, fp16)
    const int gw = warp - kGeluWarp0;
    const int q = warp & 3;    // TMEM lane quarter this warp may access
    const int half = gw >> 2;  // which 32 of the chunk's 64 columns
    const int r = q * 32 + lane;
    int gc = 0;
    for (int tl = 0; tl < n_my; ++tl) {
      for (int c = 0; c < F::NCH; ++c, ++gc) {
        const int hb = gc & 1, n = gc >> 1;
        ptx::mbar_wait(&h_full[hb], n & 1);
        ptx::tc_fence_after();
        uint32_t rr[32];
        ptx::tmem_ld_32x32(tmem_base + F::TM_H + hb * HC + (static_cast<uint32_t>(q * 32) << 16) + half * 32, rr);
        const __half2* bb = reinterpret_cast<const __half2*>(b1h + c * HC + half * 32);
        uint4 bqs[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) bqs[j] = *reinterpret_cast<const uint4*>(bb + 4 * j);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&h_empty[hb]);  // the H buffer may be overwritten by MMA1 of chunk gc+2
        ptx::mbar_wait(&g_empty[hb], (n & 1) ^ 1);      // MMA2 of chunk gc-2 has read this G buffer
        uint8_t* Gs = Gbuf(hb);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __half2* bh = reinterpret_cast<const __half2*>(&bqs[j]);
          uint4 pk;
          pk.x = gelu2_half2(__uint_as_float(rr[8 * j + 0]), __uint_as_float(rr[8 * j + 1]), bh[0]);
          pk.y = gelu2_half2(__uint_as_float(rr[8 * j + 2]), __uint_as_float(rr[8 * j + 3]), bh[1]);
          pk.z = gelu2_half2(__uint_as_float(rr[8 * j + 4]), __uint_as_float(rr[8 * j + 5]), bh[2]);
          pk.w = gelu2_half2(__uint_as_float(rr[8 * j + 6]), __uint_as_float(rr[8 * j + 7]), bh[3]);
          *reinterpret_cast<uint4*>(Gs + sw128_off(r, (half * 4 + j) * 8)) = pk;
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&g_full[hb]);
      }
    }
  } else if (warp < kMma2Warp) {
    // ================================================================== out: O (TMEM) + y -> global
    ptx::pdl_wait();  // reads y and writes out: both belong to the predecessor until it has completed
    const int q = warp & 3;
    const int l8r = lane >> 3, l8c = lane & 7;  // row segment ownership after the transposition
    float4* stg = reinterpret_cast<float4*>(smem + F::OFF_STG + (warp - kOutWarp0) * 4096);
    for (int tl = 0; tl < n_my; ++tl) {
      const int tile = first + tl * stride;
      const int ob = tl & 1;
      const int row0 = tile * TM + q * 32;
      const int nlive = p.M - row0;  // rows of this warp inside the tensor
      const long long base = static_cast<long long>(row0) * C;
      ptx::mbar_wait(&o_full[ob], (tl >> 1) & 1);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < C / 32; ++cc) {
        const int col = cc * 32 + 4 * l8c;
        // residual rows of this chunk are requested before the accumulator read
        float4 yv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int rr_ = 4 * j + l8r;
          yv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rr_ < nlive) yv[j] = *reinterpret_cast<const float4*>(p.y + base + static_cast<long long>(rr_) * C + col);
        }
        uint32_t rr[32];
        ptx::tmem_ld_32x32(tmem_base + F::TM_O + ob * C + (static_cast<uint32_t>(q * 32) << 16) + cc * 32, rr);
        ptx::tmem_ld_wait();
        if (cc == C / 32 - 1) {  // O may be overwritten by the MMA2 of tile tl+2
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&o_empty[ob]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          stg[lane * 8 + (j ^ (lane & 7))] =
              make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]),
                          __uint_as_float(rr[4 * j + 3]));
        }
        __syncwarp();
        const float4 b2 = *reinterpret_cast<const float4*>(b2g + col);
        const float4 gf = *reinterpret_cast<const float4*>(fgs + col);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int rr_ = 4 * j + l8r;
          const float4 o = stg[rr_ * 8 + (l8c ^ (rr_ & 7))];
          if (rr_ < nlive) {
            // O = (0.5 W2)(2 gelu) = W2 gelu: out = y + ffn_gamma*b2 + ffn_gamma * O
            float4 v;
            v.x = fmaf(gf.x, o.x, yv[j].x + b2.x);
            v.y = fmaf(gf.y, o.y, yv[j].y + b2.y);
            v.z = fmaf(gf.z, o.z, yv[j].z + b2.z);
            v.w = fmaf(gf.w, o.w, yv[j].w + b2.w);
            const long long off = base + static_cast<long long>(rr_) * C + col;
            *reinterpret_cast<float4*>(p.out + off) = v;
            if (p.out_bf16 != nullptr) {
              uint2 pk;
              pk.x = bf2(v.x, v.y);
              pk.y = bf2(v.z, v.w);
              *reinterpret_cast<uint2*>(p.out_bf16 + off) = pk;
            }
          }
        }
        __syncwarp();
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<512>(tmem_base);
}

bool map2d(CUtensorMap* out, const void* ptr, uint64_t cols, uint64_t rows, uint32_t box_rows) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t str[1] = {cols * 2};
  const uint32_t box[2] = {64, box_rows};
  return tmap_tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, ptr, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace

cudaError_t ffn_fused(cudaStream_t st, const bf16* a, const float* y, long long M, int C, const bf16* w1, const float* b1,
                      const void* w2_f16, const float* b2, const float* ffn_gamma, float* out, bf16* out_bf16) {
  if (C != 128 || M < 1 || M > 0x7fffffffLL / C) return cudaErrorInvalidValue;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  using F = FF<128>;
  static PerDeviceOnce once;
  {
    const cudaError_t e = once.run([] {
      return cudaFuncSetAttribute(ffn_fused_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, F::SMEM);
    });
    if (e != cudaSuccess) return e;
  }
  CUtensorMap mA, m1, m2;
  if (!map2d(&mA, a, C, static_cast<uint64_t>(M), TM)) return cudaErrorInvalidValue;  // A [M, C]: K atoms of 128 rows
  if (!map2d(&m1, w1, C, 4 * C, HC)) return cudaErrorInvalidValue;                    // W1 [4C, C]: 64 hidden rows x 64 k
  if (!map2d(&m2, w2_f16, 4 * C, C, C)) return cudaErrorInvalidValue;                  // W2 [C, 4C]: C rows x 64 k
  FfnParams p;
  p.y = y; p.out = out; p.out_bf16 = out_bf16; p.M = static_cast<int>(M);
  p.b1 = b1; p.b2 = b2; p.ffn_gamma = ffn_gamma;
  const int ntiles = static_cast<int>((M + TM - 1) / TM);
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  const cudaError_t le = launch_k(ffn_fused_kernel<128>, dim3(grid), dim3(kThreadsFfn), F::SMEM, st, mA, m1, m2, p);
  count_launch();
  return le != cudaSuccess ? le : cudaGetLastError();
}

}  // namespace stts
