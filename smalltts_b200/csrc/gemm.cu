// tcgen05 + TMA multi-tap GEMM (see gemm.cuh).  One 128 x BN output tile per CTA.
//   warp 0   : TMA producer (one lane): A box [128 rows x 64 k] and W box [BN rows x 64 k] per stage
//   warp 1   : TMEM allocator + tcgen05.mma issuer (one lane), fp32 accumulator in TMEM
//   warps 2-5: epilogue, one TMEM lane quarter each: tcgen05.ld -> bias/activation/mask/gate/residual -> global
// The epilogue is compiled per activation and only handles full, 16-byte aligned 32-column chunks (the host
// checks N % 32 == 0 and the alignments): an earlier all-in-one epilogue was ~100 KB of SASS per kernel and ran
// instruction-fetch bound (ncu: stall_no_instruction dominant).
#include "gemm.cuh"

#include <cuda_fp16.h>

#include <stdlib.h>

#include <mutex>

#include "launch.cuh"
#include "ptx.cuh"

namespace stts {

std::atomic<unsigned long long> g_launch_count{0};
thread_local int tl_capturing = 0;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int kEpiWarps = 8;
constexpr int kThreads = (2 + kEpiWarps) * 32;
constexpr int kActGelu2 = 100;  // kernel-internal: ACT_GELU with GemmEpi::gelu2_f16 (own instantiation, no general path)
constexpr int kStageBytes = 4096;  // per epilogue warp: 32 rows x 32 fp32, XOR-swizzled (no padding)

// CTAS = 2: a CTA pair (cluster of two CTAs on the SMs of one TPC) computes a 256 x BN tile with cta_group::2 MMAs; each
// CTA stages its own 128 rows of A and BN/2 rows of W, so one SM ingests 16 KB + BN*64 B per k-iteration instead of
// 16 KB + BN*128 B for the same 128 x BN x 64 MMA work (the per-SM L2 pull rate is what limits the large GEMMs).
template <int BN, int CTAS = 1>
struct Cfg {
  static constexpr int kBRows = BN / CTAS;  // W rows staged by one CTA
  static constexpr int kMaxStages = (kBRows == 256) ? 4 : (kBRows == 128 ? 6 : 8);  // one persistent CTA per SM
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = kBRows * BK * 2;
  static constexpr int smem_bytes(int stages) {
    // <= 12 barriers + slot, then 4 KB of transposition staging per epilogue warp
    return stages * (kABytes + kBBytes) + 1024 /*align slack*/ + 256 /*barriers*/ + kEpiWarps * kStageBytes;
  }
};

// Branch-free activations built on ex2/rcp (MUFU), accurate far below the bf16 rounding of their consumers.
__device__ __forceinline__ float fast_sigmoid(float x) {
  return __fdividef(1.0f, 1.0f + exp2f(-1.4426950408889634f * x));
}
// erf-GELU: 0.5 x (1 + tanh(u)) = x * sigmoid(2u) with u = x (a + b x^2 + c x^4), coefficients fitted to
// 0.5 x (1 + erf(x / sqrt 2)) on [-10, 10]: max abs error 2.6e-5 (the usual 2-term tanh form has 4.7e-4).
// Input clamped to +-8 where the quintic is still monotone; beyond that GELU(x) = max(x, 0) to fp32 accuracy.
__device__ __forceinline__ float fast_gelu(float x) {
  const float xc = fminf(fmaxf(x, -8.0f), 8.0f);
  const float x2 = xc * xc;
  // -2 * log2(e) * (a, b, c)
  const float p = fmaf(x2, fmaf(x2, 1.0153833e-3f, -0.10678167f), -2.3011139f);
  return __fdividef(x, 1.0f + exp2f(xc * p));
}
// Mish (dit.py:226-229): x * tanh(softplus(x)) = x * (w^2 + 2w) / (w^2 + 2w + 2) with w = e^x; for x > 20 -> x.
__device__ __forceinline__ float fast_mish(float x) {
  const float w = exp2f(1.4426950408889634f * fminf(x, 20.0f));
  const float n = w * (w + 2.0f);
  return x * __fdividef(n, n + 2.0f);
}

// 2*gelu on two values in packed fp16 (HFMA2 pipe + MUFU.TANH): x (1 + tanh(x (A + B x^2))), (A, B) fitted to the erf
// form (max abs error 2.7e-4, below the fp16 resolution of the result); monotone cubic, so no clamp: x^2 overflowing to
// +inf saturates tanh to +-1.  Result as fp16x2 bits.
__device__ __forceinline__ uint32_t gelu2_half2(float a, float b) {
  const __half2 x = __floats2half2_rn(a, b);
  const __half2 p = __hfma2(__hmul2(x, x), __float2half2_rn(0.03470089f), __float2half2_rn(0.80015708f));
  const __half2 u = __hmul2(x, p);
  uint32_t ui = *reinterpret_cast<const uint32_t*>(&u), ti;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(ti) : "r"(ui));
  const __half2 t = *reinterpret_cast<const __half2*>(&ti);
  const __half2 g = __hfma2(x, t, x);
  return *reinterpret_cast<const uint32_t*>(&g);
}

template <int ACT>
__device__ __forceinline__ float act_apply(float v) {
  if (ACT == ACT_GELU) return fast_gelu(v);
  if (ACT == ACT_MISH) return fast_mish(v);
  return v;
}

__device__ __forceinline__ void add4(float* v, const float* __restrict__ p) {
  const float4 x = __ldg(reinterpret_cast<const float4*>(p));
  v[0] += x.x; v[1] += x.y; v[2] += x.z; v[3] += x.w;
}
__device__ __forceinline__ void mul4(float* v, const float* __restrict__ p) {
  const float4 x = __ldg(reinterpret_cast<const float4*>(p));
  v[0] *= x.x; v[1] *= x.y; v[2] *= x.z; v[3] *= x.w;
}
__device__ __forceinline__ uint32_t bf2(float a, float b) {
  return op16_pack2(a, b);
}

// Persistent CTA: walks output tiles (tile = blockIdx.x + i * gridDim.x; n fastest, so concurrently running CTAs share
// A rows in L2).  The smem ring runs ahead across tile boundaries and the accumulator is double-buffered in TMEM
// (2 x BN columns), so the epilogue of tile i overlaps the loads and MMAs of tile i+1.
template <int BN, int ACT, int CTAS = 1, bool SPLIT = false>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmShape s,
            const GemmEpi e, const int kStages, const int n_tiles, const int m_tiles, const int total_tiles) {
  using C = Cfg<BN, CTAS>;
  constexpr int BMS = BM * CTAS;  // rows of the (pair's) tile
  // pair mode: rank 0 is the leader (issues the MMAs, owns the barriers the issuer waits on); both CTAs walk the
  // same tile sequence, CTA `rank` owns rows [rank*128, +128) of each 256-row tile and W rows [rank*BN/2, +BN/2)
  const uint32_t rank = CTAS == 2 ? ptx::cluster_ctarank() : 0u;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* smA = smem;
  uint8_t* smB = smem + kStages * C::kABytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smB + kStages * C::kBBytes);
  uint64_t* empty = full + kStages;
  uint64_t* acc_full = empty + kStages;   // [2]
  uint64_t* acc_empty = acc_full + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  uint8_t* stage_base = reinterpret_cast<uint8_t*>(full) + 256;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_t = (s.T + BMS - 1) / BMS;
  const int kchunks = (s.K + BK - 1) / BK;
  const int iters = s.taps * kchunks;
  const int first = blockIdx.x / CTAS, stride = gridDim.x / CTAS;
  // work item = (tile, part of the reduction); parts of one tile are adjacent items, i.e. run side by side
  const int S = SPLIT ? s.splits : 1;  // compile-time 1 for the regular instantiations: their code is unchanged
  const int total_items = total_tiles * S;
  const int n_my = first < total_items ? (total_items - first + stride - 1) / stride : 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&acc_full[i], 1);
      ptx::mbar_init(&acc_empty[i], kEpiWarps * CTAS);  // pair: the epilogue warps of both CTAs release the leader's copy
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CTAS == 2) ptx::tmem_alloc_pair<2 * BN>(tmem_slot);
    else ptx::tmem_alloc<2 * BN>(tmem_slot);
  }
  ptx::tc_fence_before();
  if constexpr (CTAS == 2) ptx::cluster_sync();  // the peer's barriers must be initialised before anything signals them
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the predecessor's tail.
  // Weights never depend on the predecessor, so the producer also requests the W boxes of its first stages before
  // waiting; everything else that touches global memory (A operand, residual, outputs) comes after the wait.
  ptx::pdl_trigger();
  int w_pre = 0;  // stages of the first tile whose W box is already in flight (producer warp only)
  if (CTAS == 1 && warp == 0 && n_my > 0) {
    const int tile = first / S, part = first % S;
    const int it0 = part * iters / S, it1 = (part + 1) * iters / S;
    w_pre = it1 - it0 < kStages ? it1 - it0 : kStages;
    if (ptx::elect_one()) {
      const int nx = tile % n_tiles, g = tile / (n_tiles * m_tiles);
      for (int i = 0; i < w_pre; ++i) {
        ptx::mbar_expect_tx(&full[i], C::kABytes + C::kBBytes);
        ptx::tma_load_2d(smB + i * C::kBBytes, &tmW, &full[i], (it0 + i) * BK, g * s.w_group_rows + nx * BN);
      }
    }
    __syncwarp();
  }
  ptx::pdl_wait();

  // Producer and MMA issuer are single threads running a dependent chain per k-iteration, so that chain is kept
  // free of integer divisions and descriptor rebuilds (ring position / phase / tap counters are carried, the UMMA
  // descriptors advance by a constant): measured ~690 clk per iteration before, independent of tile width and depth.
  if (warp == 0) {
    // The whole warp walks the loop (so ring state lives in uniform registers and the TMA / MMA instructions need no
    // per-lane election loops); one elected lane issues.
    uint32_t st = 0, ph = 0;  // ring position and phase, continue across tiles
    for (int li = 0; li < n_my; ++li) {
      const int item = first + li * stride;
      const int tile = item / S, part = item % S;
      const int it0 = part * iters / S, it1 = (part + 1) * iters / S;
      const int nx = tile % n_tiles, my = (tile / n_tiles) % m_tiles, g = tile / (n_tiles * m_tiles);
      const int b = my / tiles_t, t0 = (my % tiles_t) * BMS + static_cast<int>(rank) * BM, n0 = nx * BN;
      const int a_col0 = g * s.a_group_koff, w_row = g * s.w_group_rows + n0 + static_cast<int>(rank) * C::kBRows;
      int kc = it0 % kchunks, a_row = t0 + s.tap_shift0 + (it0 / kchunks) * s.tap_step;
      for (int it = it0; it < it1; ++it) {
        ptx::mbar_wait(&empty[st], ph ^ 1);
        const bool w_done = li == 0 && it - it0 < w_pre;  // expect_tx + W box already issued ahead of the PDL wait
        if (ptx::elect_one()) {
          if constexpr (CTAS == 2) {
            // the leader's barrier counts the bytes of both CTAs; the peer's copies may complete before the leader
            // arrives (the phase cannot flip until it does)
            if (rank == 0) ptx::mbar_expect_tx(&full[st], 2 * (C::kABytes + C::kBBytes));
            ptx::tma_load_3d_pair(smA + st * C::kABytes, &tmA, &full[st], a_col0 + kc * BK, a_row, b);
            ptx::tma_load_2d_pair(smB + st * C::kBBytes, &tmW, &full[st], it * BK, w_row);
          } else {
            if (!w_done) ptx::mbar_expect_tx(&full[st], C::kABytes + C::kBBytes);
            ptx::tma_load_3d(smA + st * C::kABytes, &tmA, &full[st], a_col0 + kc * BK, a_row, b);
            if (!w_done) ptx::tma_load_2d(smB + st * C::kBBytes, &tmW, &full[st], it * BK, w_row);
          }
        }
        __syncwarp();
        if (++kc == kchunks) { kc = 0; a_row += s.tap_step; }
        if (++st == static_cast<uint32_t>(kStages)) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (CTAS == 2 && rank != 0) {
      // the peer's MMA warp only allocates / frees TMEM; the leader issues for both
    } else {
    // fp16 operands: clear the bf16 format bits of A and B
    const uint32_t idesc = ptx::umma_idesc_bf16(BMS, BN) & (s.ab_f16 ? ~((1u << 7) | (1u << 10)) : ~0u);
    const uint64_t da0 = ptx::umma_desc_sw128(ptx::smem_u32(smA));
    const uint64_t db0 = ptx::umma_desc_sw128(ptx::smem_u32(smB));
    uint32_t st = 0, ph = 0;
    for (int li = 0; li < n_my; ++li) {
      const int ab = li & 1;
      // the epilogue warps must have drained this accumulator buffer (tile li-2)
      ptx::mbar_wait(&acc_empty[ab], ((li >> 1) & 1) ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + ab * BN;
      const int item = first + li * stride;
      const int it0 = (item % S) * iters / S, it1 = (item % S + 1) * iters / S;
      for (int it = it0; it < it1; ++it) {
        ptx::mbar_wait(&full[st], ph);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          // descriptor start-address field is (addr >> 4): stages are kABytes / kBBytes apart
          const uint64_t da = da0 + static_cast<uint64_t>(st * (C::kABytes >> 4));
          const uint64_t db = db0 + static_cast<uint64_t>(st * (C::kBBytes >> 4));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in the (addr >> 4) field
            if constexpr (CTAS == 2) ptx::umma_bf16_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (it != it0 || k != 0) ? 1u : 0u);
            else ptx::umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (it != it0 || k != 0) ? 1u : 0u);
          }
          // frees this smem stage (in both CTAs of a pair) once the MMAs above have read it
          if constexpr (CTAS == 2) ptx::umma_commit_pair(&empty[st]);
          else ptx::umma_commit(&empty[st]);
        }
        __syncwarp();
        if (++st == static_cast<uint32_t>(kStages)) { st = 0; ph ^= 1; }
      }
      if (ptx::elect_one()) {  // accumulator complete (pair: in both CTAs' TMEM, signalled to both epilogues)
        if constexpr (CTAS == 2) ptx::umma_commit_pair(&acc_full[ab]);
        else ptx::umma_commit(&acc_full[ab]);
      }
      __syncwarp();
    }
    }
  } else {
    // ---------------- epilogue: 8 warps; warp w owns TMEM lanes [32(w%4), +32) = tile rows and every second
    // 32-column chunk of the accumulator (chunk parity = (w-2)/4).
    // tcgen05.ld hands every lane one ROW of the chunk; storing that directly would make each store instruction
    // touch 32 different lines.  The chunk is therefore transposed through a warp-private shared-memory buffer so
    // that in the second half 8 (or 4) consecutive lanes own one row segment: bias / activation / mask / scale /
    // gate / residual and the stores then run on whole 128-byte (64-byte) row segments.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    float* stg = reinterpret_cast<float*>(stage_base + (warp - 2) * kStageBytes);
    const int l8r = lane >> 3, l8c = lane & 7;  // 32-column chunks: rows 4j + l8r (j < 8), float4 column l8c
    const int l4r = lane >> 2, l4c = lane & 3;  // 64-byte row segments: rows 8j + l4r (j < 4), 16-byte column l4c
    for (int li = 0; li < n_my; ++li) {
      const int item = first + li * stride;
      const int tile = item / S, part = item % S;
      const int nx = tile % n_tiles, my = (tile / n_tiles) % m_tiles, g = tile / (n_tiles * m_tiles);
      const int b = my / tiles_t, t0 = (my % tiles_t) * BMS + static_cast<int>(rank) * BM, n0 = nx * BN;
      const int ab = li & 1;
      const int t = t0 + row;
      const bool row_ok = t < s.T;
      const int m0 = b * s.T + t0 + q * 32;  // flattened row of this warp's first TMEM lane
      // batch index / in-batch row used by masking and gating (flattened inputs carry rows_per_batch)
      const int mrow = m0 + lane;
      const int eb_own = e.rows_per_batch > 0 ? mrow / e.rows_per_batch : b;
      const int et_own = e.rows_per_batch > 0 ? mrow % e.rows_per_batch : t;
      const bool masked = (e.row_len != nullptr) && row_ok && (et_own >= e.row_len[eb_own]);
      const uint32_t okbits = __ballot_sync(0xffffffffu, row_ok);
      const uint32_t mkbits = __ballot_sync(0xffffffffu, masked);
      const int gcol_base = g * s.out_group_cols;
      const int res_goff = g * (s.res_group_cols - s.out_group_cols);

      // post-activation part on a float4 of row rr (warp-relative), absolute output column col
      auto post_store = [&](float4 v, int rr, int col) {
        const bool mk = (mkbits >> rr) & 1u;
        if (mk && !e.mask_bf16_only) v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e.colscale != nullptr) {
          const float4 x = __ldg(reinterpret_cast<const float4*>(e.colscale + col));
          v.x *= x.x; v.y *= x.y; v.z *= x.z; v.w *= x.w;
        }
        const long long mj = m0 + rr;
        if (e.rowgate != nullptr) {
          const int ebj = e.rows_per_batch > 0 ? (m0 + rr) / e.rows_per_batch : b;
          const float4 x = __ldg(reinterpret_cast<const float4*>(e.rowgate + static_cast<long long>(ebj) * e.ld_gate + col));
          v.x *= x.x; v.y *= x.y; v.z *= x.z; v.w *= x.w;
        }
        if (e.residual != nullptr) {
          const float4 x = *reinterpret_cast<const float4*>(e.residual + mj * e.ld_res + col + res_goff);
          v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
        }
        if (e.out_f32 != nullptr) *reinterpret_cast<float4*>(e.out_f32 + mj * e.ld_out + col) = v;
        if (e.out_bf16 != nullptr) {
          const float bz = (mk && e.mask_bf16_only) ? 0.0f : 1.0f;  // bf16-only masking
          uint2 pk;
          pk.x = bf2(bz * v.x, bz * v.y);
          pk.y = bf2(bz * v.z, bz * v.w);
          *reinterpret_cast<uint2*>(e.out_bf16 + mj * e.ld_out + col) = pk;
        }
      };

      // Residual rows of the general path are fetched one chunk ahead (the first chunk of a tile before the wait
      // for its accumulator): they do not depend on the MMAs and their latency would otherwise sit between the
      // accumulator read and the store of every chunk.
      constexpr bool kGeneral = (ACT != ACT_SWIGLU16);
      const bool pre_res = kGeneral && e.residual != nullptr && !(ACT == kActGelu2);
      float4 rq[8];
      auto load_res = [&](int c, float4 (&dst)[8]) {
        const float* base = e.residual + static_cast<long long>(m0 + l8r) * e.ld_res + gcol_base + n0 + c * 32 + 4 * l8c + res_goff;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          dst[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if ((okbits >> (4 * j + l8r)) & 1u) dst[j] = *reinterpret_cast<const float4*>(base + static_cast<long long>(j) * 4 * e.ld_res);
        }
      };
      if (pre_res && n0 + half * 32 < s.N) load_res(half, rq);
      ptx::mbar_wait(&acc_full[ab], (li >> 1) & 1);
      ptx::tc_fence_after();

      // Split-K: park this part's accumulator chunks, count the arrival; only the last part of the tile goes on.
      // Scratch: [(tile, part)][chunk][8 float4 columns][128 rows] -- every warp access is 512 contiguous bytes.
      bool owner = true;
      if (SPLIT) {
#pragma unroll 1
        for (int c = half; c < BN / 32; c += 2) {
          if (n0 + c * 32 >= s.N) break;
          uint32_t r[32];
          ptx::tmem_ld_32x32(tmem_base + ab * BN + (static_cast<uint32_t>(q * 32) << 16) + c * 32, r);
          ptx::tmem_ld_wait();
          // [float4 column i][row]: the 32 lanes of a store instruction write 512 contiguous bytes
          uint4* dst = reinterpret_cast<uint4*>(e.split_scratch + ((static_cast<long long>(tile) * S + part) * (BN / 32) + c) * (BM * 32)) + row;
#pragma unroll
          for (int i = 0; i < 8; ++i) dst[i * BM] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
        }
        __threadfence();
        __syncwarp();
        int old = 0;
        if (lane == 0) {
          int* cnt = e.split_counters + static_cast<long long>(tile) * kEpiWarps + (warp - 2);
          old = atomicAdd(cnt, 1);
          if (old == S - 1) *cnt = 0;  // every part has arrived: leave the counter clean for the next launch
        }
        owner = __shfl_sync(0xffffffffu, old, 0) == S - 1;
        if (owner) __threadfence();
      }

#pragma unroll 1
      for (int c = half; owner && c < BN / 32; c += 2) {
        const int nl0 = n0 + c * 32;  // column inside the group
        if (nl0 >= s.N) break;        // warp-uniform
        const int gc0 = gcol_base + nl0;
        float4 bv[8];  // bias of the row-form paths, requested before the accumulator read
        if ((ACT == ACT_SWIGLU16 || (ACT == kActGelu2)) && e.bias != nullptr) {
#pragma unroll
          for (int i = 0; i < 8; ++i) bv[i] = __ldg(reinterpret_cast<const float4*>(e.bias + gc0 + 4 * i));
        }
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + ab * BN + (static_cast<uint32_t>(q * 32) << 16) + c * 32, r);
        ptx::tmem_ld_wait();
        if (okbits == 0u) continue;  // warp-uniform: no live row in this quarter
        if (SPLIT) {  // the tile's accumulator = sum of its parts in part order (this part's share comes from TMEM)
          const float4* src0 = reinterpret_cast<const float4*>(e.split_scratch + (static_cast<long long>(tile) * S * (BN / 32) + c) * (BM * 32)) + row;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int pp = 0; pp < S; ++pp) {
              float4 v;
              if (pp == part) {
                v = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
              } else {
                v = __ldcg(src0 + static_cast<long long>(pp) * (BN / 32) * (BM * 32 / 4) + i * BM);
              }
              if (pp == 0) acc4 = v;
              else { acc4.x += v.x; acc4.y += v.y; acc4.z += v.z; acc4.w += v.w; }
            }
            r[4 * i] = __float_as_uint(acc4.x); r[4 * i + 1] = __float_as_uint(acc4.y);
            r[4 * i + 2] = __float_as_uint(acc4.z); r[4 * i + 3] = __float_as_uint(acc4.w);
          }
        }
        if (ACT == kActGelu2) {
          // FFN hidden activation of the vocoder: 2*gelu in packed fp16, stored as fp16 (no mask/scale/residual here)
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if (e.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              v[4 * i] += bv[i].x; v[4 * i + 1] += bv[i].y; v[4 * i + 2] += bv[i].z; v[4 * i + 3] += bv[i].w;
            }
          }
          uint4* srow = reinterpret_cast<uint4*>(stg) + lane * 4;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 pk;
            pk.x = gelu2_half2(v[8 * i + 0], v[8 * i + 1]);
            pk.y = gelu2_half2(v[8 * i + 2], v[8 * i + 3]);
            pk.z = gelu2_half2(v[8 * i + 4], v[8 * i + 5]);
            pk.w = gelu2_half2(v[8 * i + 6], v[8 * i + 7]);
            srow[i ^ ((lane >> 1) & 3)] = pk;
          }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int rr = 8 * j + l4r;
            const uint4 pk = reinterpret_cast<const uint4*>(stg)[rr * 4 + (l4c ^ ((rr >> 1) & 3))];
            if ((okbits >> rr) & 1u) {
              *reinterpret_cast<uint4*>(e.out_bf16 + static_cast<long long>(m0 + rr) * e.ld_out + gc0 + 8 * l4c) = pk;
            }
          }
          __syncwarp();
          continue;
        }
        if (ACT == ACT_SWIGLU16) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if (e.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              v[4 * i] += bv[i].x; v[4 * i + 1] += bv[i].y; v[4 * i + 2] += bv[i].z; v[4 * i + 3] += bv[i].w;
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float a = v[i];
            v[i] = a * fast_sigmoid(a) * v[16 + i];
          }
          float4* srow = reinterpret_cast<float4*>(stg) + lane * 4;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            srow[i ^ ((lane >> 1) & 3)] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          __syncwarp();
          const int col = (gc0 >> 1) + 4 * l4c;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int rr = 8 * j + l4r;
            const float4 x = reinterpret_cast<const float4*>(stg)[rr * 4 + (l4c ^ ((rr >> 1) & 3))];
            if ((okbits >> rr) & 1u) post_store(x, rr, col);
          }
          __syncwarp();
          continue;
        }
        // general path: stage the raw accumulators, everything else happens on row segments
        {
          uint4* srow = reinterpret_cast<uint4*>(stg) + lane * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            srow[i ^ (lane & 7)] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
          }
        }
        __syncwarp();
        {
          const int col = gc0 + 4 * l8c;
          // phase 1: every load of the chunk is in flight before the first store
          float4 x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int rr = 4 * j + l8r;
            x[j] = reinterpret_cast<const float4*>(stg)[rr * 8 + (l8c ^ (rr & 7))];
          }
          float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), cs4 = make_float4(1.f, 1.f, 1.f, 1.f);
          if (e.bias != nullptr) bias4 = __ldg(reinterpret_cast<const float4*>(e.bias + col));
          if (e.colscale != nullptr) cs4 = __ldg(reinterpret_cast<const float4*>(e.colscale + col));
          float4 rn[8];
          const bool more = pre_res && (n0 + (c + 2) * 32 < s.N) && (c + 2 < BN / 32);
          if (more) load_res(c + 2, rn);
          // phase 2: arithmetic + stores on 128-byte row segments
          const long long obase = static_cast<long long>(m0 + l8r) * e.ld_out + col;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int rr = 4 * j + l8r;
            if (!((okbits >> rr) & 1u)) continue;
            float4 v = x[j];
            v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
            if (ACT != ACT_NONE) {
              v.x = act_apply<ACT>(v.x); v.y = act_apply<ACT>(v.y); v.z = act_apply<ACT>(v.z); v.w = act_apply<ACT>(v.w);
            }
            const bool mk = (mkbits >> rr) & 1u;
            if (mk && !e.mask_bf16_only) v = make_float4(0.f, 0.f, 0.f, 0.f);
            v.x *= cs4.x; v.y *= cs4.y; v.z *= cs4.z; v.w *= cs4.w;
            if (e.rowgate != nullptr) {  // read-only path: free to be scheduled above the stores
              const int ebj = e.rows_per_batch > 0 ? (m0 + rr) / e.rows_per_batch : b;
              const float4 gt = __ldg(reinterpret_cast<const float4*>(e.rowgate + static_cast<long long>(ebj) * e.ld_gate + col));
              v.x *= gt.x; v.y *= gt.y; v.z *= gt.z; v.w *= gt.w;
            }
            if (e.residual != nullptr) {
              v.x += rq[j].x; v.y += rq[j].y; v.z += rq[j].z; v.w += rq[j].w;
            }
            const long long off = obase + static_cast<long long>(j) * 4 * e.ld_out;
            if (e.out_f32 != nullptr) *reinterpret_cast<float4*>(e.out_f32 + off) = v;
            if (e.out_bf16 != nullptr) {
              const float bz = (mk && e.mask_bf16_only) ? 0.0f : 1.0f;  // bf16-only masking
              uint2 pk;
              pk.x = bf2(bz * v.x, bz * v.y);
              pk.y = bf2(bz * v.z, bz * v.w);
              *reinterpret_cast<uint2*>(e.out_bf16 + off) = pk;
            }
          }
          if (more) {
#pragma unroll
            for (int j = 0; j < 8; ++j) rq[j] = rn[j];
          }
        }
        __syncwarp();
      }
      // this warp is done reading the accumulator buffer
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CTAS == 2) ptx::mbar_arrive_leader(&acc_empty[ab]);
        else ptx::mbar_arrive(&acc_empty[ab]);
      }
    }
  }

  ptx::tc_fence_before();
  // pair: neither CTA may free its TMEM / leave while the other still reads its shared memory or signals its barriers
  if constexpr (CTAS == 2) ptx::cluster_sync();
  else __syncthreads();
  if (warp == 1) {
    if constexpr (CTAS == 2) ptx::tmem_dealloc_pair<2 * BN>(tmem_base);
    else ptx::tmem_dealloc<2 * BN>(tmem_base);
  }
}

// ------------------------------------------------------------------ host side
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode_fn() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeFn>(p);
    }
  });
  return fn;
}

bool make_tmap(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box) {
  EncodeFn fn = get_encode_fn();
  if (fn == nullptr) return false;
  cuuint64_t gdim[3];
  cuuint64_t gstr[2];
  cuuint32_t bx[3];
  cuuint32_t es[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BN, int ACT, int CTAS = 1, bool SPLIT = false>
cudaError_t launch_inst(cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmShape& s,
                        const GemmEpi& e) {
  using C = Cfg<BN, CTAS>;
  static PerDeviceOnce attr_set;
  {
    const cudaError_t err = attr_set.run([] {
      return cudaFuncSetAttribute(gemm_kernel<BN, ACT, CTAS, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C::smem_bytes(C::kMaxStages));
    });
    if (err != cudaSuccess) return err;
  }
  const int tiles_t = (s.T + BM * CTAS - 1) / (BM * CTAS);
  const int n_tiles = (s.N + BN - 1) / BN, m_tiles = s.B * tiles_t;
  const long long total = static_cast<long long>(n_tiles) * m_tiles * s.groups;
  if (total > 0x7fffffffLL) return cudaErrorInvalidValue;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int slots = num_sms / CTAS;  // persistent CTAs (CTA pairs)
  const int splits = SPLIT ? s.splits : 1;
  const long long items = total * splits;
  if (items > 0x7fffffffLL) return cudaErrorInvalidValue;
  const int grid = items < slots ? static_cast<int>(items) : slots;
  // deep ring (the producer runs ahead into the next tile); short reductions of small problems need fewer stages
  const int iters = s.taps * ((s.K + BK - 1) / BK) / splits;
  const long long ring = static_cast<long long>(iters) * ((items + grid - 1) / grid);
  int stages = ring < 2 ? 2 : (ring > C::kMaxStages ? C::kMaxStages : static_cast<int>(ring));
  {
    static int forced = -1;  // STTS_GEMM_STAGES=n: pipeline-depth experiments (tools/bench_gemm.py)
    if (forced < 0) {
      const char* ev = getenv("STTS_GEMM_STAGES");
      forced = ev ? atoi(ev) : 0;
    }
    if (forced >= 2 && forced < stages) stages = forced;
  }
  const cudaError_t le =
      CTAS == 2 ? launch_k_pair(gemm_kernel<BN, ACT, CTAS, SPLIT>, dim3(2 * grid), dim3(kThreads), C::smem_bytes(stages), stream, tmA,
                                tmW, s, e, stages, n_tiles, m_tiles, static_cast<int>(total))
                : launch_k(gemm_kernel<BN, ACT, CTAS, SPLIT>, dim3(grid), dim3(kThreads), C::smem_bytes(stages), stream, tmA, tmW,
                           s, e, stages, n_tiles, m_tiles, static_cast<int>(total));
  count_launch();
  return le != cudaSuccess ? le : cudaGetLastError();
}

// CTA-pair instantiations exist for the epilogues of the large vocoder GEMMs only (ACT_NONE and the fp16 2*gelu form).
template <int BN>
cudaError_t launch_pair(cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmShape& s,
                        const GemmEpi& e) {
  if (e.act == ACT_NONE) return launch_inst<BN, ACT_NONE, 2>(stream, tmA, tmW, s, e);
  if (e.act == ACT_GELU && e.gelu2_f16) return launch_inst<BN, kActGelu2, 2>(stream, tmA, tmW, s, e);
  return cudaErrorInvalidValue;
}

// Split-K instantiations exist for the epilogues of the vocoder's wave-quantised GEMMs (plain and fp16 2*gelu).
template <int BN>
cudaError_t launch_split(cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmShape& s,
                         const GemmEpi& e) {
  if (e.act == ACT_NONE) return launch_inst<BN, ACT_NONE, 1, true>(stream, tmA, tmW, s, e);
  if (e.act == ACT_GELU && e.gelu2_f16) return launch_inst<BN, kActGelu2, 1, true>(stream, tmA, tmW, s, e);
  return cudaErrorInvalidValue;
}

template <int BN>
cudaError_t launch_bn(cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmShape& s,
                      const GemmEpi& e) {
  if (s.splits > 1) {
    if (BN == 128 || BN == 256) return launch_split<(BN == 128 || BN == 256) ? BN : 128>(stream, tmA, tmW, s, e);
    return cudaErrorInvalidValue;
  }
  switch (e.act) {
    case ACT_NONE:
      return launch_inst<BN, ACT_NONE>(stream, tmA, tmW, s, e);
    case ACT_GELU:
      if (e.gelu2_f16) return launch_inst<BN, kActGelu2>(stream, tmA, tmW, s, e);
      return launch_inst<BN, ACT_GELU>(stream, tmA, tmW, s, e);
    case ACT_SWIGLU16:
      return launch_inst<BN, ACT_SWIGLU16>(stream, tmA, tmW, s, e);
    case ACT_MISH:
      if (BN == 64) return launch_inst<64, ACT_MISH>(stream, tmA, tmW, s, e);
      return cudaErrorInvalidValue;
    default:
      return cudaErrorInvalidValue;
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

long long gemm_split_counters(const GemmShape& s, int block_n) {
  block_n &= ~kGemmPairFlag;
  const long long tiles = static_cast<long long>(s.B) * ((s.T + BM - 1) / BM) * ((s.N + block_n - 1) / block_n) * s.groups;
  return tiles * kEpiWarps;
}
long long gemm_split_scratch_floats(const GemmShape& s, int block_n) {
  block_n &= ~kGemmPairFlag;
  return gemm_split_counters(s, block_n) / kEpiWarps * (s.splits > 1 ? s.splits : 1) * BM * block_n;
}

cudaError_t launch_gemm(cudaStream_t stream, int block_n, const GemmA& a, const GemmW& w, const GemmShape& s,
                        const GemmEpi& e) {
  if (s.T <= 0 || s.B <= 0 || s.N <= 0 || s.K <= 0) return cudaErrorInvalidValue;
  if (s.groups > 1 && (s.a_group_koff % 8) != 0) return cudaErrorInvalidValue;  // TMA: 16-byte aligned box start
  if ((a.ld % 8) != 0 || (w.ld % 8) != 0 || !aligned16(a.ptr) || !aligned16(w.ptr)) return cudaErrorInvalidValue;
  // the epilogue works on full, 16-byte aligned 32-column chunks
  if ((s.N % 32) != 0 || (s.out_group_cols % 32) != 0 || (s.res_group_cols % 4) != 0) return cudaErrorInvalidValue;
  const int ld_req = e.out_bf16 != nullptr ? 8 : 4;
  if ((e.ld_out % ld_req) != 0 || !aligned16(e.out_f32) || !aligned16(e.out_bf16)) return cudaErrorInvalidValue;
  if (e.residual != nullptr && ((e.ld_res % 4) != 0 || !aligned16(e.residual))) return cudaErrorInvalidValue;
  if (!aligned16(e.bias) || !aligned16(e.colscale) || !aligned16(e.rowgate) || (e.ld_gate % 4) != 0) {
    return cudaErrorInvalidValue;
  }
  // CTA-pair variant: requested explicitly (block_n | kGemmPairFlag: tests, micro-benchmarks) or, with
  // STTS_GEMM_2CTA=1, wherever it measured faster in isolation (tools/bench_gemm.py, profiles/r01_gemm_microbench_pair.txt):
  // reductions of >= 16 k-iterations (K*taps >= 1024: -4..-12 %; shorter ones are epilogue-bound and lose), same tile
  // width as the single-CTA choice, enough tiles to occupy all 74 pairs.
  bool pair = (block_n & kGemmPairFlag) != 0;
  block_n &= ~kGemmPairFlag;
  if (s.splits > 1) {
    const int iters = s.taps * ((s.K + BK - 1) / BK);
    if (pair || s.splits > 8 || iters < s.splits || e.split_scratch == nullptr || e.split_counters == nullptr ||
        !aligned16(e.split_scratch)) {
      return cudaErrorInvalidValue;
    }
  }
  if (!pair && (block_n == 128 || block_n == 256) && (e.act == ACT_NONE || (e.act == ACT_GELU && e.gelu2_f16))) {
    static int env = -1;
    if (env < 0) {
      const char* ev = getenv("STTS_GEMM_2CTA");
      env = (ev && ev[0] == '1') ? 1 : 0;
    }
    const long long tiles2 = static_cast<long long>(s.B) * ((s.T + 255) / 256) * ((s.N + block_n - 1) / block_n) * s.groups;
    pair = env == 1 && s.taps * ((s.K + BK - 1) / BK) >= 16 && tiles2 >= 60;
  }
  if (pair && block_n != 128 && block_n != 256) return cudaErrorInvalidValue;
  CUtensorMap tmA, tmW;
  {
    uint64_t dims[3] = {static_cast<uint64_t>(a.cols), static_cast<uint64_t>(s.T), static_cast<uint64_t>(s.B)};
    uint64_t str[2] = {static_cast<uint64_t>(a.ld) * 2, static_cast<uint64_t>(a.ld) * 2 * static_cast<uint64_t>(s.T)};
    uint32_t box[3] = {BK, BM, 1};
    if (!make_tmap(&tmA, a.ptr, 3, dims, str, box)) return cudaErrorInvalidValue;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(w.ld), static_cast<uint64_t>(w.rows)};
    uint64_t str[1] = {static_cast<uint64_t>(w.ld) * 2};
    uint32_t box[2] = {BK, static_cast<uint32_t>(pair ? block_n / 2 : block_n)};
    if (!make_tmap(&tmW, w.ptr, 2, dims, str, box)) return cudaErrorInvalidValue;
  }
  if (pair) return block_n == 256 ? launch_pair<256>(stream, tmA, tmW, s, e) : launch_pair<128>(stream, tmA, tmW, s, e);
  switch (block_n) {
    case 32:
      return launch_bn<32>(stream, tmA, tmW, s, e);
    case 64:
      return launch_bn<64>(stream, tmA, tmW, s, e);
    case 128:
      return launch_bn<128>(stream, tmA, tmW, s, e);
    case 256:
      return launch_bn<256>(stream, tmA, tmW, s, e);
    default:
      return cudaErrorInvalidValue;
  }
}

}  // namespace stts
