// Chained DiT GEMMs on tcgen05 (see dit_chain.cuh for what is fused and why).
//
// Persistent CTAs (one per SM) pull 128 x bn output tiles from ONE global queue: tiles are numbered phase by phase
// (phase = one GEMM of the chain) and claimed in that order with an atomic counter.  A tile only depends on tiles with
// smaller numbers, and a claimed tile is always in the hands of a resident CTA, so the kernel makes progress with ANY
// number of resident CTAs (two engines sharing a GPU cannot deadlock each other) and the tail of one phase overlaps the
// head of the next.  The claiming warp hands the tile numbers to the other roles through a small shared-memory FIFO:
//   warp 0      W producer : claims tiles (one ahead), TMA boxes of the weight tile (no dependency: runs ahead through
//                            the ring, also across phase boundaries and ahead of the PDL wait)
//   warp 1      A producer : waits until the row block it needs has been completed by the previous phase (a counter in
//                            global memory, bumped by the epilogue warps of the producing tiles), then TMA-loads A
//   warp 2      MMA issuer : tcgen05.mma 128 x bn x 16, fp32 accumulators double-buffered in TMEM (2 x 256 columns)
//   warps 4-11  epilogue   : tcgen05.ld -> LayerNorm fold / head RMSNorm + RoPE / SwiGLU / gated residual + LayerNorm
//                            partials -> global; then release the accumulator and bump the row block's counter
// The tile width is a per-tile value (64 / 128 / 192): it only changes the instruction descriptor, the number of
// 64-row weight boxes per stage and the epilogue's chunk count, so one launch mixes widths to keep every phase at or
// below one tile per SM at the benchmark shape (q|k|v: 24 head tiles of 128 + 5 gate tiles of 192 per row block).
//
// Memory model notes: data that another CTA wrote earlier in the SAME launch is read either by TMA (A operands; the
// A producer acquires the counter at gpu scope and crosses to the async proxy with fence.proxy.async) or with
// ld.global.cg (residual rows, LayerNorm partials: L1 may hold lines from an earlier phase).  Write-after-read hazards
// between phases are excluded by the chain itself: a buffer's next writer depends transitively on all of its readers.
#include "dit_chain.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

#include "chain_attn.cuh"
#include "launch.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace stts {

namespace {

constexpr int BM = 128, BK = 64;
constexpr int kStages = 4;
constexpr int kABytes = BM * BK * 2;    // 16 KB: one k-block (64 columns) of the A tile
// One ring slot = 48 KB.  Wide tiles (bn 128 / 192) put one k-block in it (A 16 KB | W <= 24 KB); the narrow tiles of
// to_out / w2 / velocity (bn 64) put TWO (A0 | A1 | W0 8 KB | W1 8 KB): the mainloop is paced by the bytes in flight per
// SM (ring bytes / L2 latency), so the 24 KB stages of a narrow tile would leave half of the ring empty.
constexpr int kSlotBytes = 48 * 1024;
constexpr int kEpiWarps = 8;
constexpr int kEpiWarp0 = 4;
constexpr int kThreads = (kEpiWarp0 + kEpiWarps) * 32;
constexpr int kStgBytes = 2048;         // per epilogue warp: 32 rows x 16 fp32 (or 32 bf16), XOR-swizzled
constexpr int kAccCols = 256;           // TMEM columns per accumulator buffer
constexpr int kTq = 4;                  // depth of the tile FIFO between the claiming warp and the other roles
constexpr int kTqReaders = 2 + kEpiWarps;  // A producer, MMA issuer, epilogue warps
constexpr int kColvecFloats = 192;         // per epilogue warp: 3 chunks x (cs | b') or 2 chunks x (cs | b') + norm weights
constexpr int kCtrlBytes = 1024;            // mbarriers, tile FIFO, TMEM slot, phase table
constexpr int kSmem = kStages * kSlotBytes + kCtrlBytes + kEpiWarps * kStgBytes + 2 * 2 * BM * 4 + kEpiWarps * kColvecFloats * 4 + 1024;
static_assert(kSmem <= 232448, "chain kernel does not fit in shared memory");
static_assert(chain_attn::kSmemBytes <= kStages * kSlotBytes, "an attention item borrows the operand ring");

struct ChainMaps {
  CUtensorMap a[3];  // xb [M, 960], ob [M, 1024], hb [M, 2400]: box 64 x 128 rows
  CUtensorMap w[5];  // per ChainKind, all blocks stacked: box 64 x 64 rows
  chain_attn::Maps at;
};

struct ChainDev {
  ChainWeights w;
  ChainBuffers b;
  ChainCall c;
  int m_tiles;
  // CHAIN_ATTN phases: items (utterance, 64-query tile, head), head fastest; to_out of row block m waits for
  // attn_target[m] items (those whose query rows touch the block)
  int q_tiles, attn_items;
  int attn_target[kChainMaxRowBlocks];
};

// ---------------------------------------------------------------- static shape tables
__device__ __forceinline__ int kind_ntiles(int kind) {
  return kind == CHAIN_QKVG ? 29 : (kind == CHAIN_W13 ? 25 : (kind == CHAIN_VEL ? 1 : 15));
}
__device__ __forceinline__ int kind_kiters(int kind) {
  return kind == CHAIN_OUT ? 16 : (kind == CHAIN_W2 ? 38 : 15);
}
__device__ __forceinline__ int kind_wrows(int kind) {  // weight rows per block
  return kind == CHAIN_QKVG ? kChainQKVG : (kind == CHAIN_W13 ? kChainW13 : (kind == CHAIN_VEL ? 0 : kChainD));
}
__device__ __forceinline__ int kind_amap(int kind) { return kind == CHAIN_OUT ? 1 : (kind == CHAIN_W2 ? 2 : 0); }
__device__ __forceinline__ void tile_cols(int kind, int n, int& n0, int& bn) {
  if (kind == CHAIN_QKVG) {
    if (n < 24) { n0 = n * 128; bn = 128; } else { n0 = 3072 + (n - 24) * 192; bn = 192; }
  } else if (kind == CHAIN_W13) {
    n0 = n * 192; bn = 192;
  } else {
    n0 = n * 64; bn = 64;
  }
}

#ifndef STTS_CHAIN_KPB
#define STTS_CHAIN_KPB 2  // k-blocks per ring slot for the narrow (bn 64) tiles; 1 = A/B experiments
#endif
__device__ __forceinline__ int slot_kblocks(int bn) { return bn == 64 ? STTS_CHAIN_KPB : 1; }

struct Tile {
  int p, kind, blk, m, n, g;  // attention items: m = utterance * q_tiles + query tile, n = head
};
__device__ __forceinline__ int phase_items(const ChainDev& d, int p) {
  return d.c.kind[p] == CHAIN_ATTN ? d.attn_items : kind_ntiles(d.c.kind[p]) * d.m_tiles;
}
// Global tile number -> (phase, row block, column tile); phases are numbered back to back (pstart: first tile of every
// phase, in shared memory), n fastest inside a phase.  A role sees increasing tile numbers: `cur` is its phase cursor.
__device__ __forceinline__ Tile decode_tile(int g, const ChainDev& d, const int* pstart, int& cur) {
  while (g >= pstart[cur + 1]) ++cur;
  Tile t;
  t.g = g;
  t.p = cur;
  g -= pstart[cur];
  t.kind = d.c.kind[cur];
  t.blk = d.c.blk[cur];
  const int nt = t.kind == CHAIN_ATTN ? kChainH : kind_ntiles(t.kind);
  t.m = g / nt;
  t.n = g % nt;
  return t;
}
// completed items of phase p a consumer of row block m waits for
__device__ __forceinline__ int phase_target(const ChainDev& d, int p, int m) {
  return d.c.kind[p] == CHAIN_ATTN ? d.attn_target[m] : kind_ntiles(d.c.kind[p]);
}

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Wait until *p >= target.  Polls with relaxed loads (an acquire load per poll would invalidate this SM's L1 every time,
// under the feet of the epilogue warps) and backs off, then acquires once.  Bounded like ptx::mbar_wait: a scheduling
// bug traps instead of hanging the GPU.
__device__ __forceinline__ void wait_flag(const int* p, int target) {
  uint32_t spins = 0, ns = 64;
  while (ld_relaxed(p) < target) {
    __nanosleep(ns);
    if (ns < 512) ns += ns;
    if (++spins > (1u << 23)) __trap();
  }
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}

// Optional role timeline (ChainBuffers::trace != nullptr, tools/trace_chain.py): trace[(cta * 64 + seq) * 16 + ev] =
// globaltimer ns; ev 0 tile published | 1 A producer at the flag | 2 flag seen | 3 first operands landed | 4 MMAs issued
// | 5 epilogue sees the accumulator | 6 epilogue stores issued | 7 tile counted; slot 15 = global tile number + 1.
__device__ __forceinline__ void trace_ev(unsigned long long* trace, int seq, int ev, long long val = -1) {
  if (trace == nullptr || seq >= 64) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  trace[(static_cast<size_t>(blockIdx.x) * 64 + seq) * 16 + ev] = val >= 0 ? static_cast<unsigned long long>(val) : t;
}

__device__ __forceinline__ uint32_t bf2(float a, float b) {
  return op16_pack2(a, b);
}
// 2^x on the special-function unit, no range fix-ups (no branches): inputs here are <= 0 or moderate, tiny results flush to 0
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_sigmoid(float x) {
  return __fdividef(1.0f, 1.0f + exp2f(-1.4426950408889634f * x));
}

// ---------------------------------------------------------------- epilogue building blocks
struct Epi {
  int q, half, lane;    // TMEM lane quarter, chunk parity of this warp, lane
  float* stg;           // warp-private staging (4 KB)
  float* ssx;           // [2][2][128] head RMSNorm exchange between the two warps of a lane quarter
  uint32_t acc;         // TMEM address of this tile's accumulator at this warp's lane quarter
  int m0;               // global row of this warp's first TMEM lane
  uint32_t okbits;      // rows of this warp inside [0, M)
};

// mean / rstd of LayerNorm(x[row]) from the 30 partials the producing epilogue left (dit.py:16, eps 1e-6)
__device__ __forceinline__ void row_stats(const float* __restrict__ stats, int grow, bool ok, float& mean, float& rstd) {
  float s1 = 0.f, s2 = 0.f;
  if (ok) {
    const float4* p = reinterpret_cast<const float4*>(stats + static_cast<long long>(grow) * (2 * kChainParts));
#pragma unroll
    for (int i = 0; i < kChainParts / 2; ++i) {
      const float4 v = __ldcg(p + i);
      s1 += v.x + v.z;
      s2 += v.y + v.w;
    }
  }
  mean = s1 * (1.0f / kChainD);
  const float var = fmaxf(s2 * (1.0f / kChainD) - mean * mean, 0.f);
  rstd = rsqrtf(var + 1e-6f);
}

// v = rstd * (acc - mean * cs) + b  on one 32-column chunk held in row form
__device__ __forceinline__ void ln_chunk(const uint32_t (&r)[32], float (&v)[32], float mean, float rstd,
                                         const float* __restrict__ cs, const float* __restrict__ bb) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 c4 = reinterpret_cast<const float4*>(cs)[i];  // shared memory, same address in every lane: broadcast
    const float4 b4 = reinterpret_cast<const float4*>(bb)[i];
    v[4 * i + 0] = fmaf(rstd, fmaf(-mean, c4.x, __uint_as_float(r[4 * i + 0])), b4.x);
    v[4 * i + 1] = fmaf(rstd, fmaf(-mean, c4.y, __uint_as_float(r[4 * i + 1])), b4.y);
    v[4 * i + 2] = fmaf(rstd, fmaf(-mean, c4.z, __uint_as_float(r[4 * i + 2])), b4.z);
    v[4 * i + 3] = fmaf(rstd, fmaf(-mean, c4.w, __uint_as_float(r[4 * i + 3])), b4.w);
  }
}

// 32 rows x 32 bf16 (row form) -> global rows of pitch ld (elements), 64-byte row segments per 4 lanes
__device__ __forceinline__ void store_bf16_chunk(const Epi& e, const float (&v)[32], bf16* __restrict__ out, long long ld,
                                                 int col0) {
  uint4* srow = reinterpret_cast<uint4*>(e.stg) + e.lane * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 pk;
    pk.x = bf2(v[8 * i + 0], v[8 * i + 1]);
    pk.y = bf2(v[8 * i + 2], v[8 * i + 3]);
    pk.z = bf2(v[8 * i + 4], v[8 * i + 5]);
    pk.w = bf2(v[8 * i + 6], v[8 * i + 7]);
    srow[i ^ ((e.lane >> 1) & 3)] = pk;
  }
  __syncwarp();
  const int l4r = e.lane >> 2, l4c = e.lane & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int rr = 8 * j + l4r;
    const uint4 pk = reinterpret_cast<const uint4*>(e.stg)[rr * 4 + (l4c ^ ((rr >> 1) & 3))];
    if ((e.okbits >> rr) & 1u) {
      *reinterpret_cast<uint4*>(out + static_cast<long long>(e.m0 + rr) * ld + col0 + 8 * l4c) = pk;
    }
  }
  __syncwarp();
}

// 32 rows x 32 fp32 (row form) -> global: two passes of 16 columns, 64-byte row segments per 4 lanes
__device__ __forceinline__ void store_f32_chunk(const Epi& e, const float (&v)[32], float* __restrict__ out, long long ld,
                                                int col0) {
  const int l4r = e.lane >> 2, l4c = e.lane & 3;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float4* srow = reinterpret_cast<float4*>(e.stg) + e.lane * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      srow[i ^ ((e.lane >> 1) & 3)] = make_float4(v[16 * h + 4 * i], v[16 * h + 4 * i + 1], v[16 * h + 4 * i + 2], v[16 * h + 4 * i + 3]);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rr = 8 * j + l4r;
      const float4 x = reinterpret_cast<const float4*>(e.stg)[rr * 4 + (l4c ^ ((rr >> 1) & 3))];
      if ((e.okbits >> rr) & 1u) {
        *reinterpret_cast<float4*>(out + static_cast<long long>(e.m0 + rr) * ld + col0 + 16 * h + 4 * l4c) = x;
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------- attention item, epilogue-warp side (chain_attn.cuh)
// Softmax over the score accumulator and the gated output.  The two warps of a lane quarter split the 32-key chunks of S by
// parity and exchange row max / row sum through `ssx` (named barrier 1 + quarter, like the head RMSNorm of the q|k|v tiles).
// Not inlined: the GEMM epilogues around the call site must not pay for its registers.  Returns the updated exchange count.
__device__ __noinline__ int attn_epilogue(const ChainDev& d, Epi e, uint32_t tmem_base, uint8_t* ring, uint64_t* s_full,
                                          uint64_t* p_full, uint64_t* o_full, uint32_t par, int hcount, int b, int h, int q0,
                                          unsigned long long* trow) {
  const ChainCall& c = d.c;
  const chain_attn::Keys ak = chain_attn::keys_of(c, b);
  const int lane = e.lane;
  const int row = e.q * 32 + lane;  // row of the item = TMEM lane
  const bool row_ok = q0 + row < c.T;
  e.m0 = b * c.T + q0 + e.q * 32;
  e.okbits = __ballot_sync(0xffffffffu, row_ok);
  const bool live = e.okbits != 0u;  // a quarter past the end of the utterance only keeps the barriers company
  const uint32_t s_acc = tmem_base + chain_attn::kSCol + (static_cast<uint32_t>(e.q * 32) << 16);
  const uint32_t o_acc = tmem_base + chain_attn::kOCol + (static_cast<uint32_t>(e.q * 32) << 16);
  const float scale_log2 = 1.4426950408889634f * 0.09128709291752769f;  // log2(e) / sqrt(120)
  const int nch = (ak.e2 + 31) >> 5;  // 32-key chunks of S that hold keys
  auto stamp = [&](int ev) {
    if (trow != nullptr && e.q == 0 && e.half == 0 && lane == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      trow[ev] = t;
    }
  };

  // ---- sigmoid(gate) of the item's rows -> shared memory (fp16, [128 rows][128 dims], 16-byte chunks XOR-swizzled by the
  // row), fetched while Q / K are still in flight.  Coalesced: a warp reads one row (120 floats) per instruction -- the
  // per-thread row-form read (32 rows x 16 bytes per instruction) is bound by the load/store unit, not by L2.
  // fp16: values in (0, 1), 2^-12 absolute error, well below the rounding of the 16-bit output they multiply.
  uint8_t* sgate = ring + chain_attn::kGateOff;
  {
    const int wid = e.half * 4 + e.q;  // 0..7: rows 16 wid .. + 16
    const int rows_live = min(c.T - q0, chain_attn::kRows);
    constexpr float kNegLog2e = -1.4426950408889634f;
    float4 g4[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int rr = wid * 16 + i;
      g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rr < rows_live && lane < kChainHD / 4) {  // written in this launch by other CTAs: L2, not L1
        g4[i] = __ldcg(reinterpret_cast<const float4*>(d.b.gate + static_cast<long long>(b * c.T + q0 + rr) * kChainD + h * kChainHD) + lane);
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int rr = wid * 16 + i;
      const __half2 lo = __floats2half2_rn(__fdividef(1.0f, 1.0f + ex2_approx(kNegLog2e * g4[i].x)),
                                           __fdividef(1.0f, 1.0f + ex2_approx(kNegLog2e * g4[i].y)));
      const __half2 hi = __floats2half2_rn(__fdividef(1.0f, 1.0f + ex2_approx(kNegLog2e * g4[i].z)),
                                           __fdividef(1.0f, 1.0f + ex2_approx(kNegLog2e * g4[i].w)));
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&lo);
      pk.y = *reinterpret_cast<const uint32_t*>(&hi);
      // dims 4 lane .. + 4 = 8 bytes at byte 8 lane of the row: 16-byte chunk lane / 2, half lane & 1
      *reinterpret_cast<uint2*>(sgate + rr * 256 + ((((lane >> 1) ^ (rr & 15)) << 4) | ((lane & 1) << 3))) = pk;
    }
  }
  ptx::named_bar_sync(6, kEpiWarps * 32);  // (also orders the previous item's reads of the gate tile before these writes)
  ptx::mbar_wait(s_full, par);
  ptx::tc_fence_after();
#ifndef STTS_ATTN_DEBUG_TRACE
  stamp(8);
#endif
  // ---- pass 1: row max over this warp's chunks (padded keys masked)
  float mx = -INFINITY;
  if (live) {
#pragma unroll 1
    for (int cc = e.half; cc < nch; cc += 2) {
      const uint32_t vm = __ballot_sync(0xffffffffu, chain_attn::key_valid(ak, cc * 32 + lane));
      uint32_t r[32];
      ptx::tmem_ld_32x32(s_acc + cc * 32, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if ((vm >> i) & 1u) mx = fmaxf(mx, __uint_as_float(r[i]));
      }
    }
  }
  {
    float* sx = e.ssx + (hcount & 1) * (2 * BM);
    sx[e.half * BM + row] = mx;
    ptx::named_bar_sync(1 + e.q, 64);
    mx = fmaxf(sx[row], sx[BM + row]);
    ++hcount;
  }
  // ---- pass 2: P = exp2((S - max) * scale) as bf16 into the A-operand layout (K-major, 128-byte swizzle, 64-key blocks)
  float sum = 0.f;
  if (live) {
    const float mxs = mx == -INFINITY ? 0.f : mx * scale_log2;
#pragma unroll 1
    for (int cc = e.half; cc < nch; cc += 2) {
      const uint32_t vm = __ballot_sync(0xffffffffu, chain_attn::key_valid(ak, cc * 32 + lane));
      uint32_t r[32];
      ptx::tmem_ld_32x32(s_acc + cc * 32, r);
      ptx::tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {  // masked keys: 2^-inf = 0 (selects, no branches around the special-function unit)
        const float s0 = (vm >> (2 * i)) & 1u ? __uint_as_float(r[2 * i]) : -INFINITY;
        const float s1 = (vm >> (2 * i + 1)) & 1u ? __uint_as_float(r[2 * i + 1]) : -INFINITY;
        const float p0 = ex2_approx(fmaf(s0, scale_log2, -mxs));
        const float p1 = ex2_approx(fmaf(s1, scale_log2, -mxs));
        sum += p0 + p1;
        pk[i] = op16_pack2(p0, p1);
      }
      uint8_t* prow = ring + chain_attn::kPOff + (cc >> 1) * (chain_attn::kRows * 128) + row * 128;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int ch = (cc & 1) * 4 + q4;
        *reinterpret_cast<uint4*>(prow + ((ch ^ (row & 7)) << 4)) = make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
      }
    }
  }
  fence_proxy_async_all();  // P (generic-proxy writes to shared memory) -> the MMA's async-proxy reads
  ptx::tc_fence_before();
  __syncwarp();
  if (lane == 0) ptx::mbar_arrive(p_full);
#ifdef STTS_ATTN_DEBUG_TRACE
  if (trow != nullptr && lane == 0 && e.half * 4 + e.q < 6) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trow[8 + e.half * 4 + e.q] = t;
  }
#else
  stamp(9);
#endif
  {
    float* sx = e.ssx + (hcount & 1) * (2 * BM);
    sx[e.half * BM + row] = sum;
    ptx::named_bar_sync(1 + e.q, 64);
    sum = sx[row] + sx[BM + row];
    ++hcount;
  }
  const float inv = sum > 0.f ? 1.0f / sum : 0.f;
  ptx::mbar_wait(o_full, par);
  ptx::tc_fence_after();
#ifndef STTS_ATTN_DEBUG_TRACE
  stamp(10);
#else
  stamp(14);
#endif
  if (live) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int cc = e.half + 2 * j;
      uint32_t r[32];
      float v[32];
      ptx::tmem_ld_32x32(o_acc + cc * 32, r);
      uint2 sg[8];  // this row's gates of dims 32 cc .. + 32: four 16-byte chunks
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const uint4 x = *reinterpret_cast<const uint4*>(sgate + row * 256 + (((cc * 4 + q4) ^ (row & 15)) << 4));
        sg[2 * q4] = make_uint2(x.x, x.y);
        sg[2 * q4 + 1] = make_uint2(x.z, x.w);
      }
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 g01 = __half22float2(*reinterpret_cast<const __half2*>(&sg[i].x));
        const float2 g23 = __half22float2(*reinterpret_cast<const __half2*>(&sg[i].y));
        // the pad dims (120..127) hold exact zeros: V's pad columns are zero
        v[4 * i + 0] = __uint_as_float(r[4 * i + 0]) * inv * g01.x;
        v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) * inv * g01.y;
        v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) * inv * g23.x;
        v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) * inv * g23.y;
      }
      store_bf16_chunk(e, v, d.b.ob + h * kChainHDP, kChainH * kChainHDP, cc * 32);
    }
  }
  ptx::tc_fence_before();
#ifndef STTS_ATTN_DEBUG_TRACE
  stamp(11);
#endif
  return hcount;
}

// ---------------------------------------------------------------- the kernel
// kAttn: instantiation that can run CHAIN_ATTN phases (the attention item costs registers around its call site; launches
// without such a phase use the lean instantiation)
template <bool kAttn>
__global__ void __launch_bounds__(kThreads, 1)
dit_chain_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainDev d) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* ring = smem;  // kStages slots of kSlotBytes
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + kStages * kSlotBytes);
  uint64_t* empty = full + kStages;
  uint64_t* acc_full = empty + kStages;  // [2]
  uint64_t* acc_empty = acc_full + 2;    // [2]
  uint64_t* tq_full = acc_empty + 2;     // [kTq]
  uint64_t* tq_empty = tq_full + kTq;    // [kTq]
  uint64_t* dep_full = tq_empty + kTq;   // [kTq] the A producer has seen the row block's flag for the tile in this FIFO slot
  uint64_t* attn_bars = dep_full + kTq;  // [chain_attn::kBarriers] attention item: qk_full, v_full, s_full, p_full, o_full
  uint64_t* const at_qk_full = attn_bars, * const at_v_full = attn_bars + 1, * const at_s_full = attn_bars + 2,
                * const at_p_full = attn_bars + 3, * const at_o_full = attn_bars + 4;
  uint64_t* attn_done = attn_bars + chain_attn::kBarriers;  // an attention item has released the ring
  int* tileq = reinterpret_cast<int*>(attn_done + 1);   // [kTq]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tileq + kTq);
  int* pstart = reinterpret_cast<int*>(tmem_slot + 1);  // [n_phases + 1]
  static_assert((3 * kStages + 4 + 3 * kTq + chain_attn::kBarriers + 1) * 8 + kTq * 4 + 4 + (kChainMaxPhases + 1) * 4 <= kCtrlBytes,
                "control block overflows");
  uint8_t* stage_base = reinterpret_cast<uint8_t*>(full) + kCtrlBytes;
  float* ssx = reinterpret_cast<float*>(stage_base + kEpiWarps * kStgBytes);
  float* colvec = ssx + 2 * 2 * BM;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const ChainCall& c = d.c;
  const int m_tiles = d.m_tiles;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 3; ++i) ptx::prefetch_tmap(&maps.a[i]);
    for (int i = 0; i < 5; ++i) ptx::prefetch_tmap(&maps.w[i]);
    ptx::mbar_init(at_qk_full, 2);  // A producer (Q) + W producer (K), each with its own byte count
    ptx::mbar_init(at_v_full, 1);
    ptx::mbar_init(at_s_full, 1);
    ptx::mbar_init(at_p_full, kEpiWarps);
    ptx::mbar_init(at_o_full, 1);
    ptx::mbar_init(attn_done, 1);
    if (kAttn) {
      ptx::prefetch_tmap(&maps.at.q);
      ptx::prefetch_tmap(&maps.at.kv_self[0]);
      ptx::prefetch_tmap(&maps.at.kv_self[1]);
      ptx::prefetch_tmap(&maps.at.ref);
      ptx::prefetch_tmap(&maps.at.text);
    }
    pstart[0] = 0;
    for (int p = 0; p < c.n_phases; ++p) pstart[p + 1] = pstart[p] + phase_items(d, p);
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full[i], 2);  // W producer + A producer (each arrives with its own byte count)
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&acc_full[i], 1);
      ptx::mbar_init(&acc_empty[i], kEpiWarps);
    }
    for (int i = 0; i < kTq; ++i) {
      ptx::mbar_init(&tq_full[i], 1);
      ptx::mbar_init(&tq_empty[i], kTqReaders);
      ptx::mbar_init(&dep_full[i], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<2 * kAccCols>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_trigger();

  const int total_tiles = pstart[c.n_phases];
  // counters of this launch (dit_chain.cuh chain_ready_ints): completed items per (phase, row block), then the claim
  // counter on a line of its own
  int* const done_cnt = d.b.ready;
  int* const next_tile = d.b.ready + ((c.n_phases * m_tiles + 31) & ~31) + 32;

  // Consumer side of the tile FIFO: calls fn(tile, slot, parity) for every tile this CTA claimed, in claim order.  The
  // slot is handed back only after the tile has been processed, so per-slot state (dep_full) cannot be recycled early.
  auto walk = [&](auto&& fn) {
    uint32_t slot = 0, tph = 0;
    int cur = 0;
    for (;;) {
      ptx::mbar_wait(&tq_full[slot], tph);
      const int g = tileq[slot];
      __syncwarp();
      if (g >= 0) fn(decode_tile(g, d, pstart, cur), slot, tph);
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tq_empty[slot]);
      if (g < 0) break;
      if (++slot == kTq) { slot = 0; tph ^= 1; }
    }
  };

  if (warp == 0) {
    // ================================================================== W producer + tile claims
    uint32_t st = 0, ph = 0, slot = 0, tph = 0;
    int issued = 0;
    bool waited = false;
    auto claim = [&]() {  // the whole warp gets the next global tile number
      int g = 0;
      if (lane == 0) g = atomicAdd(next_tile, 1);
      return __shfl_sync(0xffffffffu, g, 0);
    };
    int g = claim(), g_raw = 0, wseq = 0, cur = 0;
    uint32_t attn_seen = 0, last_st = 0, last_ph = 0;  // last ring slot filled (and the round it was filled in)
    for (;;) {
      const bool live = g < total_tiles;
      const uint32_t my_slot = slot, my_tph = tph;
      ptx::mbar_wait(&tq_empty[slot], tph ^ 1);
      if (lane == 0) {
        tileq[slot] = live ? g : -1;
        ptx::mbar_arrive(&tq_full[slot]);  // release: the tile number is visible to whoever sees this phase complete
      }
      __syncwarp();
      if (++slot == kTq) { slot = 0; tph ^= 1; }
      if (!live) break;
      const Tile t = decode_tile(g, d, pstart, cur);
      if (lane == 0) { trace_ev(d.b.trace, wseq, 0); trace_ev(d.b.trace, wseq, 15, g + 1); }
      ++wseq;
      if (kAttn && t.kind == CHAIN_ATTN) {
        // ---- attention item: this warp loads K and V (one lane per 16-key group and dim half) into the ring, which the
        // item borrows: first every operand of the previous tile must have been consumed
        if (!waited) {
          ptx::pdl_wait();
          waited = true;
        }
        if (lane == 0) g_raw = atomicAdd(next_tile, 1);
        const int b = t.m / d.q_tiles, h = t.n;
        const chain_attn::Keys ak = chain_attn::keys_of(c, b);
        if (issued > 0) ptx::mbar_wait(&empty[last_st], last_ph);
        if (lane == 0) {
          ptx::mbar_expect_tx(at_qk_full, static_cast<uint32_t>(ak.ng * 4096));
          ptx::mbar_expect_tx(at_v_full, static_cast<uint32_t>(ak.ng * 4096));
        }
        __syncwarp();
        const int kg = lane >> 1, hf = lane & 1, par = c.qkv_db ? (t.blk & 1) : 0;
        const bool mine = kg < ak.ng, self = kg * 16 < ak.e0;
        // the cross-attention caches do not depend on this launch: requested while the q|k|v phase may still be running
        if (mine && !self) chain_attn::load_group(maps.at, c, ak, t.blk, par, b, h, kg, hf, ring, at_qk_full, at_v_full);
        ptx::mbar_wait(&dep_full[my_slot], my_tph);  // raised by the A producer: q|k|v of the utterance are complete
        if (t.p > 0) asm volatile("fence.acq_rel.gpu;" ::: "memory");
        fence_proxy_async_all();
        if (mine && self) chain_attn::load_group(maps.at, c, ak, t.blk, par, b, h, kg, hf, ring, at_qk_full, at_v_full);
        __syncwarp();
        ptx::mbar_wait(attn_done, attn_seen & 1);  // nothing may be loaded into the ring until the item is done with it
        ++attn_seen;
        g = __shfl_sync(0xffffffffu, g_raw, 0);
        continue;
      }
      int n0, bn;
      tile_cols(t.kind, t.n, n0, bn);
      const int wrow = t.blk * kind_wrows(t.kind) + n0;
      const int kblocks = kind_kiters(t.kind), kpb = slot_kblocks(bn);
      const int iters = (kblocks + kpb - 1) / kpb;  // ring slots of this tile
      const CUtensorMap* tm = &maps.w[t.kind];
      for (int it = 0; it < iters; ++it) {
        // weights never depend on the predecessor kernel: the first ring's worth of boxes is requested before the
        // PDL wait; the wait itself must still happen before this kernel can be considered complete
        if (!waited && issued == kStages) {
          ptx::pdl_wait();
          waited = true;
        }
        // Claim the next tile late: close to when this CTA can really start it (so that tiles go to whoever is free),
        // but a ring's depth before the end of this tile's loads, which hides the round trip of the atomic.
        if (it == (iters > kStages ? iters - kStages : 0) && lane == 0) g_raw = atomicAdd(next_tile, 1);
        ptx::mbar_wait(&empty[st], ph ^ 1);
        if (ptx::elect_one()) {
          const int nkb = kblocks - it * kpb < kpb ? kblocks - it * kpb : kpb;
          uint8_t* wdst = ring + st * kSlotBytes + (kpb == 2 ? 2 * kABytes : kABytes);
          ptx::mbar_expect_tx(&full[st], static_cast<uint32_t>(bn * 128 * nkb));
          for (int kb = 0; kb < nkb; ++kb) {
            for (int r = 0; r < bn / 64; ++r) {
              ptx::tma_load_2d(wdst + kb * (bn * 128) + r * 8192, tm, &full[st], (it * kpb + kb) * BK, wrow + r * 64);
            }
          }
        }
        __syncwarp();
        ++issued;
        last_st = st;
        last_ph = ph;
        if (++st == kStages) { st = 0; ph ^= 1; }
      }
      g = __shfl_sync(0xffffffffu, g_raw, 0);  // the atomic's result is first needed here
    }
    if (!waited) ptx::pdl_wait();
  } else if (warp == 1) {
    // ================================================================== A producer
    ptx::pdl_wait();
    uint32_t st = 0, ph = 0, attn_seen = 0, last_st = 0, last_ph = 0;
    bool filled = false;
    int aseq = 0;
    walk([&](const Tile& t, uint32_t slot, uint32_t) {
      const int p = t.p, kind = t.kind, m = t.m;
      if (lane == 0) trace_ev(d.b.trace, aseq, 1);
      if (kAttn && kind == CHAIN_ATTN) {
        // q of the item's rows and k, v of the whole utterance: every row block the utterance touches
        if (p > 0 && lane == 0) {
          const int b = m / d.q_tiles;
          const int m_lo = (b * c.T) / BM, m_hi = (b * c.T + c.T - 1) / BM;
          for (int mm = m_lo; mm <= m_hi; ++mm) wait_flag(done_cnt + (p - 1) * m_tiles + mm, phase_target(d, p - 1, mm));
        }
        __syncwarp();
        fence_proxy_async_all();  // q was written through the generic proxy by other CTAs; read below by TMA
        if (lane == 0) ptx::mbar_arrive(&dep_full[slot]);
        if (lane == 0) trace_ev(d.b.trace, aseq, 2);
        ++aseq;
        if (filled) ptx::mbar_wait(&empty[last_st], last_ph);  // the previous tile's operands have been consumed
        if (ptx::elect_one()) {
          const int b = m / d.q_tiles, q0 = (m % d.q_tiles) * chain_attn::kRows, par = c.qkv_db ? (t.blk & 1) : 0;
          ptx::mbar_expect_tx(at_qk_full, chain_attn::kQBytes);
          for (int hf = 0; hf < 2; ++hf) {
            ptx::tma_load_3d(ring + chain_attn::kQOff + hf * (chain_attn::kRows * 128), &maps.at.q, at_qk_full, hf * 64, t.n,
                             par * 3 * c.M + b * c.T + q0);
          }
        }
        __syncwarp();
        ptx::mbar_wait(attn_done, attn_seen & 1);  // the ring belongs to the item until then
        ++attn_seen;
        return;
      }
      if (p > 0) {
        // all tiles of the previous phase that write rows [128 m, +128) must be done
        if (lane == 0) wait_flag(done_cnt + (p - 1) * m_tiles + m, phase_target(d, p - 1, m));
        __syncwarp();
        fence_proxy_async_all();  // generic-proxy writes of other CTAs -> this thread's async-proxy (TMA) reads
      }
      // one poller per CTA: the epilogue warps learn about the flag through shared memory
      if (lane == 0) ptx::mbar_arrive(&dep_full[slot]);
      if (lane == 0) trace_ev(d.b.trace, aseq, 2);
      ++aseq;
      int n0, bn;
      tile_cols(kind, t.n, n0, bn);
      const int kblocks = kind_kiters(kind), kpb = slot_kblocks(bn);
      const int iters = (kblocks + kpb - 1) / kpb;
      const CUtensorMap* tm = &maps.a[kind_amap(kind)];
      for (int it = 0; it < iters; ++it) {
        ptx::mbar_wait(&empty[st], ph ^ 1);
        if (ptx::elect_one()) {
          const int nkb = kblocks - it * kpb < kpb ? kblocks - it * kpb : kpb;
          ptx::mbar_expect_tx(&full[st], static_cast<uint32_t>(kABytes * nkb));
          for (int kb = 0; kb < nkb; ++kb) {
            ptx::tma_load_2d(ring + st * kSlotBytes + kb * kABytes, tm, &full[st], (it * kpb + kb) * BK, m * BM);
          }
        }
        __syncwarp();
        last_st = st;
        last_ph = ph;
        filled = true;
        if (++st == kStages) { st = 0; ph ^= 1; }
      }
    });
  } else if (warp == 2) {
    // ================================================================== MMA issuer
    ptx::pdl_wait();
    const uint64_t d0 = ptx::umma_desc_sw128(ptx::smem_u32(ring));
    uint32_t st = 0, ph = 0;
    int li = 0, tseq = 0;
    uint32_t attn_seen = 0;
    walk([&](const Tile& t, uint32_t, uint32_t) {
      const int kind = t.kind;
      if (kAttn && kind == CHAIN_ATTN) {
        // ---- attention item: S = Q K^T into TMEM columns [0, 256), then O = P V into [256, 384).  Both accumulator
        // buffers of the GEMM tiles are idle here (their epilogues ran before this item's) and li does not advance.
        const chain_attn::Keys ak = chain_attn::keys_of(c, t.m / d.q_tiles);
        const uint32_t par = attn_seen & 1;
        ++attn_seen;
        ptx::mbar_wait(at_qk_full, par);
        ptx::tc_fence_after();
        if (lane == 0) trace_ev(d.b.trace, tseq, 3);
        if (ptx::elect_one()) {
          if (ak.ng > 0) {
            const uint32_t idesc = ptx::umma_idesc_bf16(BM, static_cast<uint32_t>(ak.ng * 16));
            const uint64_t dq = d0 + (chain_attn::kQOff >> 4), dk = d0 + (chain_attn::kKOff >> 4);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                ptx::umma_bf16(tmem_base + chain_attn::kSCol, dq + kb * ((chain_attn::kRows * 128) >> 4) + 2 * k,
                               dk + kb * ((chain_attn::kMaxKeys * 128) >> 4) + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
              }
            }
          }
          ptx::umma_commit(at_s_full);
        }
        __syncwarp();
        ptx::mbar_wait(at_v_full, par);
#ifdef STTS_ATTN_DEBUG_TRACE
        if (lane == 0) trace_ev(d.b.trace, tseq, 3);
#endif
        ptx::mbar_wait(at_p_full, par);  // the epilogue warps have written P (and are done reading S)
        ptx::tc_fence_after();
#ifdef STTS_ATTN_DEBUG_TRACE
        if (lane == 0) trace_ev(d.b.trace, tseq, 1);
#endif
        if (ptx::elect_one()) {
          const uint32_t idesc = ptx::umma_idesc_bf16(BM, chain_attn::kHD) | chain_attn::kIdescBMajorMN;
          const uint64_t dp = d0 + (chain_attn::kPOff >> 4);
          const uint32_t v_addr = ptx::smem_u32(ring) + chain_attn::kVOff;
          for (int kg = 0; kg < ak.ng; ++kg) {  // K step = one 16-key group
            ptx::umma_bf16(tmem_base + chain_attn::kOCol, dp + (kg >> 2) * ((chain_attn::kRows * 128) >> 4) + 2 * (kg & 3),
                           chain_attn::umma_desc_mn_sw128(v_addr + kg * 2048, chain_attn::kMaxKeys * 128), idesc, kg != 0 ? 1u : 0u);
          }
          ptx::umma_commit(at_o_full);
        }
        __syncwarp();
        if (lane == 0) trace_ev(d.b.trace, tseq, 4);
        ++tseq;
        return;
      }
      int n0, bn;
      tile_cols(kind, t.n, n0, bn);
      const uint32_t idesc = ptx::umma_idesc_bf16(BM, static_cast<uint32_t>(bn));
      const int kblocks = kind_kiters(kind), kpb = slot_kblocks(bn);
      const int iters = (kblocks + kpb - 1) / kpb;
      const int ab = li & 1;
      ptx::mbar_wait(&acc_empty[ab], ((li >> 1) & 1) ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + ab * kAccCols;
      for (int it = 0; it < iters; ++it) {
        ptx::mbar_wait(&full[st], ph);
        ptx::tc_fence_after();
        if (it == 0 && lane == 0) trace_ev(d.b.trace, tseq, 3);
        if (ptx::elect_one()) {
          const int nkb = kblocks - it * kpb < kpb ? kblocks - it * kpb : kpb;
          // descriptor start-address field is (addr >> 4)
          const uint64_t ds = d0 + static_cast<uint64_t>(st * (kSlotBytes >> 4));
          const uint32_t w_off = (kpb == 2 ? 2 * kABytes : kABytes) >> 4;
          for (int kb = 0; kb < nkb; ++kb) {
            const uint64_t da = ds + static_cast<uint64_t>(kb * (kABytes >> 4));
            const uint64_t db = ds + w_off + static_cast<uint64_t>(kb * ((bn * 128) >> 4));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              ptx::umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (it | kb | k) != 0 ? 1u : 0u);
            }
          }
          ptx::umma_commit(&empty[st]);
        }
        __syncwarp();
        if (++st == kStages) { st = 0; ph ^= 1; }
      }
      if (ptx::elect_one()) ptx::umma_commit(&acc_full[ab]);
      __syncwarp();
      if (lane == 0) trace_ev(d.b.trace, tseq, 4);
      ++li;
      ++tseq;
    });
  } else if (warp >= kEpiWarp0) {
    // ================================================================== epilogue (8 warps)
    ptx::pdl_wait();
    Epi e;
    e.q = warp & 3;
    e.half = (warp - kEpiWarp0) >> 2;
    e.lane = lane;
    e.stg = reinterpret_cast<float*>(stage_base + (warp - kEpiWarp0) * kStgBytes);
    e.ssx = ssx;
    // warp-private column vectors of the current tile (fold vectors, head norm weights): fetched BEFORE the waits
    float* cv = colvec + (warp - kEpiWarp0) * kColvecFloats;
    int li = 0, hcount = 0, tseq = 0;
    uint32_t attn_seen = 0;
    walk([&](const Tile& t, uint32_t slot, uint32_t tph) {
      const int p = t.p, kind = t.kind, blk = t.blk, m = t.m, n = t.n;
      if constexpr (kAttn) if (kind == CHAIN_ATTN) {
        // ---- attention item (utterance b, 128 queries from q0, head n): softmax + output on these eight warps
        const int b = m / d.q_tiles, q0 = (m % d.q_tiles) * chain_attn::kRows;
        ptx::mbar_wait(&dep_full[slot], tph);
        if (p > 0) asm volatile("fence.acq_rel.gpu;" ::: "memory");  // the gate rows come from the q|k|v|gate phase
        if (warp == kEpiWarp0 && lane == 0) trace_ev(d.b.trace, tseq, 5);
        hcount = attn_epilogue(d, e, tmem_base, ring, at_s_full, at_p_full, at_o_full, attn_seen & 1, hcount, b, n, q0,
                               d.b.trace != nullptr && tseq < 64 ? d.b.trace + (static_cast<size_t>(blockIdx.x) * 64 + tseq) * 16 : nullptr);
        ++attn_seen;
        if (warp == kEpiWarp0 && lane == 0) trace_ev(d.b.trace, tseq, 6);
        fence_proxy_async_all();  // ob is read by TMA (to_out); the ring is next written by TMA
        ptx::named_bar_sync(5, kEpiWarps * 32);
        if (warp == kEpiWarp0 && lane == 0) {
          ptx::mbar_arrive(attn_done);
          __threadfence();
          const int r0 = b * c.T + q0, r1 = b * c.T + min(q0 + chain_attn::kRows, c.T) - 1;
          for (int mm = r0 / BM; mm <= r1 / BM; ++mm) atomicAdd(done_cnt + p * m_tiles + mm, 1);
          trace_ev(d.b.trace, tseq, 7);
        }
        ++tseq;
        return;
      }
      int n0, bn;
      tile_cols(kind, n, n0, bn);
      const int ab = li & 1;
      e.acc = tmem_base + ab * kAccCols + (static_cast<uint32_t>(e.q * 32) << 16);
      e.m0 = m * BM + e.q * 32;
      const int grow = e.m0 + lane;  // row this thread owns in row form
      const bool row_ok = grow < c.M;
      e.okbits = __ballot_sync(0xffffffffu, row_ok);
      const bool ln_kind = kind == CHAIN_QKVG || kind == CHAIN_W13 || kind == CHAIN_VEL;
      const bool head_tile = kind == CHAIN_QKVG && n < 24;
      const int kind3 = n >> 3, head = n & 7;  // head tiles: q (0) | k (1) | v (2)
      const int nch = kind == CHAIN_VEL ? 1 : (bn == 192 ? 3 : (bn == 128 ? 2 : 1));  // chunks of this warp

      // ---- (1) everything that does not depend on earlier phases, requested before any wait: the tile's fold vectors
      // (cs | b' per chunk) and head norm weights go to warp-private shared memory, per-lane vectors to registers.
      float4 rc4[4], rs4[4];   // RoPE cos / sin of this thread's row (head tiles, rotated chunk = chunk `half`)
      // to_out / w2 work on 64-byte row segments: lane (l4r, l4c) owns rows 8 j + l4r and, in pass h, the float4 column
      // n0 + 32 half + 16 h + 4 l4c.  Per-column vectors of both passes: bias, tanh(gate), scale of the consuming norm.
      const int l4r = lane >> 2, l4c = lane & 3;
      const int col = n0 + e.half * 32 + 4 * l4c;
      float4 bias4[2], g4[2], s4[2];
      bias4[0] = bias4[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ln_kind) {
        const float* cs = kind == CHAIN_QKVG ? c.fold + kFoldCsQ + blk * kChainQKVG + n0
                                             : (kind == CHAIN_W13 ? c.fold + kFoldCs13 + blk * kChainW13 + n0 : c.fold + kFoldCsV);
        const float* bb = kind == CHAIN_QKVG ? c.fold + kFoldBq + blk * kChainQKVG + n0
                                             : (kind == CHAIN_W13 ? c.fold + kFoldB13 + blk * kChainW13 + n0 : c.fold + kFoldBv);
        // layout: chunk j of this warp (columns 32 (half + 2 j)) -> cv[64 j .. +32) = cs, cv[64 j + 32 .. +32) = b'
        for (int i = lane; i < nch * 16; i += 32) {
          const int j = i >> 4, w = (i >> 3) & 1, f = i & 7;
          const float* src = (w ? bb : cs) + (e.half + 2 * j) * 32;
          reinterpret_cast<float4*>(cv + j * 64 + w * 32)[f] = __ldg(reinterpret_cast<const float4*>(src) + f);
        }
        if (head_tile && kind3 < 2) {
          const float* nw = (kind3 == 0 ? d.w.qn : d.w.kn) + blk * (kChainH * kChainHD) + head * kChainHD;
          if (lane < 16) {  // cv[128 + 32 j .. +32): norm weights of chunk j, zero beyond the 120 real dims
            const int j = lane >> 3, f = lane & 7, cc = e.half + 2 * j;
            float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (cc * 32 + 4 * f < kChainHD) w4 = __ldg(reinterpret_cast<const float4*>(nw + cc * 32) + f);
            reinterpret_cast<float4*>(cv + 128 + j * 32)[f] = w4;
          }
          const int pos = grow % c.T;  // rotated dims [0, 64) = chunks 0 and 1: chunk `half` of this warp
          const float4* cp = reinterpret_cast<const float4*>(d.w.cos_t + static_cast<long long>(pos) * 32 + e.half * 16);
          const float4* sp = reinterpret_cast<const float4*>(d.w.sin_t + static_cast<long long>(pos) * 32 + e.half * 16);
#pragma unroll
          for (int i = 0; i < 4; ++i) { rc4[i] = __ldg(cp + i); rs4[i] = __ldg(sp + i); }
        }
      } else {
        const bool is_out = kind == CHAIN_OUT;
        const float* mblk = c.mod + blk * 6 * kChainD;
        const float* gatev = mblk + (is_out ? 2 : 5) * kChainD;
        // scale of the LayerNorm that consumes this x: scale_mlp of this block | scale_msa of the next | final scale
        const float* scalev = is_out ? mblk + 4 * kChainD
                                     : (blk + 1 < kChainBlocks ? mblk + 7 * kChainD : c.mod + kChainBlocks * 6 * kChainD);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (!is_out) bias4[h] = __ldg(reinterpret_cast<const float4*>(d.w.b2 + blk * kChainD + col + 16 * h));
          g4[h] = __ldg(reinterpret_cast<const float4*>(gatev + col + 16 * h));
          s4[h] = __ldg(reinterpret_cast<const float4*>(scalev + col + 16 * h));
        }
      }
      __syncwarp();

      // ---- (2) what earlier phases of this launch produced (LayerNorm partials, residual rows): the row block's flag
      // first -- the MMAs of this tile cannot start before it either, so this wait is off the critical path
      ptx::mbar_wait(&dep_full[slot], tph);  // raised by the A producer once it has seen the flag (one poller per CTA)
      if (p > 0) asm volatile("fence.acq_rel.gpu;" ::: "memory");
      float mean = 0.f, rstd = 0.f;
      float4 xr[8];
      uint32_t mkbits = 0u;
      if (ln_kind) {
        row_stats(d.b.stats, grow, row_ok, mean, rstd);
      } else {
        bool masked = false;
        if (kind == CHAIN_OUT && row_ok) masked = (grow % c.T) >= __ldg(c.frames + grow / c.T);
        mkbits = __ballot_sync(0xffffffffu, masked);
#pragma unroll
        for (int i = 0; i < 8; ++i) {  // i = 4 h + j: pass h, row 8 j + l4r
          const int rr = 8 * (i & 3) + l4r;
          xr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if ((e.okbits >> rr) & 1u) {
            xr[i] = __ldcg(reinterpret_cast<const float4*>(d.b.x + static_cast<long long>(e.m0 + rr) * kChainD + col + 16 * (i >> 2)));
          }
        }
      }

      // ---- (3) the accumulator
      ptx::mbar_wait(&acc_full[ab], (li >> 1) & 1);
      ptx::tc_fence_after();
      if (warp == kEpiWarp0 && lane == 0) trace_ev(d.b.trace, tseq, 5);

      if (e.okbits != 0u) {
        if (head_tile) {
          // ---- one head of q, k or v: LayerNorm fold, per-head RMSNorm, RoPE, bf16 head layout
          float rn = 1.0f;
          if (kind3 < 2) {
            float ss = 0.f;
#pragma unroll 1
            for (int j = 0; j < 2; ++j) {
              uint32_t r[32];
              float v[32];
              ptx::tmem_ld_32x32(e.acc + (e.half + 2 * j) * 32, r);
              ptx::tmem_ld_wait();
              ln_chunk(r, v, mean, rstd, cv + j * 64, cv + j * 64 + 32);
#pragma unroll
              for (int i = 0; i < 32; ++i) ss = fmaf(v[i], v[i], ss);  // pad columns are exactly 0 (zero weights, cs, b)
            }
            float* sx = e.ssx + (hcount & 1) * (2 * BM);
            sx[e.half * BM + e.q * 32 + lane] = ss;
            ptx::named_bar_sync(1 + e.q, 64);  // the two warps of this lane quarter
            const float tot = sx[e.q * 32 + lane] + sx[BM + e.q * 32 + lane];
            rn = rsqrtf(tot * (1.0f / kChainHD) + 1e-6f);
            ++hcount;
          }
          bf16* out = d.b.qkv + static_cast<long long>(kind3 + (c.qkv_db ? 3 * (blk & 1) : 0)) * c.M * (kChainH * kChainHDP) +
                      head * kChainHDP;
#pragma unroll 1
          for (int j = 0; j < 2; ++j) {
            const int cc = e.half + 2 * j;
            uint32_t r[32];
            float v[32];
            ptx::tmem_ld_32x32(e.acc + cc * 32, r);
            ptx::tmem_ld_wait();
            ln_chunk(r, v, mean, rstd, cv + j * 64, cv + j * 64 + 32);
            if (kind3 < 2) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 w4 = reinterpret_cast<const float4*>(cv + 128 + j * 32)[i];
                v[4 * i + 0] *= rn * w4.x; v[4 * i + 1] *= rn * w4.y;
                v[4 * i + 2] *= rn * w4.z; v[4 * i + 3] *= rn * w4.w;
              }
              if (j == 0) {  // interleaved-pair rotation of dims [0, 64): pair i of chunk cc uses angle index 16 cc + i
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float cj[4] = {rc4[i].x, rc4[i].y, rc4[i].z, rc4[i].w}, sj[4] = {rs4[i].x, rs4[i].y, rs4[i].z, rs4[i].w};
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const float x0 = v[8 * i + 2 * k], x1 = v[8 * i + 2 * k + 1];
                    v[8 * i + 2 * k] = x0 * cj[k] - x1 * sj[k];
                    v[8 * i + 2 * k + 1] = x1 * cj[k] + x0 * sj[k];
                  }
                }
              }
            }
            store_bf16_chunk(e, v, out, kChainH * kChainHDP, cc * 32);
          }
        } else if (kind == CHAIN_QKVG) {
          // ---- attention gate columns (dit.py:111): LayerNorm fold only, fp32 (the attention kernel applies sigmoid)
#pragma unroll 1
          for (int j = 0; j < 3; ++j) {
            uint32_t r[32];
            float v[32];
            ptx::tmem_ld_32x32(e.acc + (e.half + 2 * j) * 32, r);
            ptx::tmem_ld_wait();
            ln_chunk(r, v, mean, rstd, cv + j * 64, cv + j * 64 + 32);
            store_f32_chunk(e, v, d.b.gate, kChainD, n0 - 3072 + (e.half + 2 * j) * 32);
          }
        } else if (kind == CHAIN_W13) {
          // ---- SwiGLU hidden (dit.py:186): chunk = 16 w1 columns | 16 w3 columns
#pragma unroll 1
          for (int j = 0; j < 3; ++j) {
            const int cc = e.half + 2 * j;
            uint32_t r[32];
            float v[32];
            ptx::tmem_ld_32x32(e.acc + cc * 32, r);
            ptx::tmem_ld_wait();
            ln_chunk(r, v, mean, rstd, cv + j * 64, cv + j * 64 + 32);
            uint4 pk[2];
            uint32_t* pw = reinterpret_cast<uint32_t*>(pk);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float a0 = v[2 * i], a1 = v[2 * i + 1];
              pw[i] = bf2(a0 * fast_sigmoid(a0) * v[16 + 2 * i], a1 * fast_sigmoid(a1) * v[16 + 2 * i + 1]);
            }
            // 32 rows x 16 bf16: 32-byte rows staged, two lanes per row store 16 bytes each
            uint4* srow = reinterpret_cast<uint4*>(e.stg) + lane * 2;
            srow[0 ^ ((lane >> 2) & 1)] = pk[0];
            srow[1 ^ ((lane >> 2) & 1)] = pk[1];
            __syncwarp();
            const int hcol = (n0 + cc * 32) >> 1;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int rr = 16 * k + (lane >> 1), hf = lane & 1;
              const uint4 x = reinterpret_cast<const uint4*>(e.stg)[rr * 2 + (hf ^ ((rr >> 2) & 1))];
              if ((e.okbits >> rr) & 1u) {
                *reinterpret_cast<uint4*>(d.b.hb + static_cast<long long>(e.m0 + rr) * kChainFF + hcol + 8 * hf) = x;
              }
            }
            __syncwarp();
          }
        } else if (kind == CHAIN_VEL) {
          // ---- velocity head (model.py:100) behind the final adaLN (dit.py:35-39): one chunk per warp
          uint32_t r[32];
          float v[32];
          ptx::tmem_ld_32x32(e.acc + e.half * 32, r);
          ptx::tmem_ld_wait();
          ln_chunk(r, v, mean, rstd, cv, cv + 32);
          store_f32_chunk(e, v, d.b.vel, 64, e.half * 32);
        } else {
          // ---- to_out / w2: x += tanh(gate) * (acc + b) with padded query rows masked for to_out (dit.py:115-118,
          // 198,201); writes x, its LayerNorm partials and the next GEMM's operand bf16(x * (1 + scale)).
          uint32_t r[32];
          ptx::tmem_ld_32x32(e.acc + e.half * 32, r);
          ptx::tmem_ld_wait();
          float p1[4] = {0.f, 0.f, 0.f, 0.f}, p2[4] = {0.f, 0.f, 0.f, 0.f};  // LayerNorm partials of rows 8 j + l4r
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            {
              uint4* srow = reinterpret_cast<uint4*>(e.stg) + lane * 4;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                srow[i ^ ((lane >> 1) & 3)] = make_uint4(r[16 * h + 4 * i], r[16 * h + 4 * i + 1], r[16 * h + 4 * i + 2], r[16 * h + 4 * i + 3]);
              }
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int rr = 8 * j + l4r;
              float4 v = reinterpret_cast<const float4*>(e.stg)[rr * 4 + (l4c ^ ((rr >> 1) & 3))];
              const float4 xv = xr[4 * h + j];
              v.x += bias4[h].x; v.y += bias4[h].y; v.z += bias4[h].z; v.w += bias4[h].w;
              if ((mkbits >> rr) & 1u) v = make_float4(0.f, 0.f, 0.f, 0.f);
              v.x = fmaf(v.x, g4[h].x, xv.x); v.y = fmaf(v.y, g4[h].y, xv.y);
              v.z = fmaf(v.z, g4[h].z, xv.z); v.w = fmaf(v.w, g4[h].w, xv.w);
              p1[j] += (v.x + v.y) + (v.z + v.w);
              p2[j] += fmaf(v.x, v.x, v.y * v.y) + fmaf(v.z, v.z, v.w * v.w);
              if ((e.okbits >> rr) & 1u) {
                const long long off = static_cast<long long>(e.m0 + rr) * kChainD + col + 16 * h;
                *reinterpret_cast<float4*>(d.b.x + off) = v;
                uint2 pk;
                pk.x = bf2(v.x * (1.0f + s4[h].x), v.y * (1.0f + s4[h].y));
                pk.y = bf2(v.z * (1.0f + s4[h].z), v.w * (1.0f + s4[h].w));
                *reinterpret_cast<uint2*>(d.b.xb + off) = pk;
              }
            }
            __syncwarp();
          }
          const int part = (n0 >> 5) + e.half;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float s1 = p1[j], s2 = p2[j];
            s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 2); s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
            const int rr = 8 * j + l4r;
            if (l4c == 0 && ((e.okbits >> rr) & 1u)) {
              *reinterpret_cast<float2*>(d.b.stats + (static_cast<long long>(e.m0 + rr) * kChainParts + part) * 2) =
                  make_float2(s1, s2);
            }
          }
        }
      } else if (head_tile && kind3 < 2) {
        // no live row in this quarter, but the partner-warp barrier of the head epilogue is unconditional
        ptx::named_bar_sync(1 + e.q, 64);
        ++hcount;
      }
      // this warp is done with the accumulator buffer ...
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc_empty[ab]);
      // ... and with its share of the tile.  Publish the tile once all eight warps are done: each thread makes its
      // global writes visible to the async proxy (the consumer reads them with TMA), the barrier orders them before
      // thread 0, whose gpu-scope fence + atomic is the (cumulative) release.
      if (warp == kEpiWarp0 && lane == 0) trace_ev(d.b.trace, tseq, 6);
      fence_proxy_async_all();
      ptx::named_bar_sync(5, kEpiWarps * 32);
      if (warp == kEpiWarp0 && lane == 0) {
        __threadfence();
        atomicAdd(done_cnt + p * m_tiles + m, 1);  // the consumers' (one per CTA) pollers watch this count
        trace_ev(d.b.trace, tseq, 7);
      }
      ++li;
      ++tseq;
    });
  } else {
    ptx::pdl_wait();  // warp 3: idle
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<2 * kAccCols>(tmem_base);
}

// ---------------------------------------------------------------- per-timestep folded vectors
// One warp per output row j of a GEMM whose input is LayerNorm(x) * (1 + scale) + shift:
//   cs[j] = sum_k W[j,k] (1 + scale[k]),   b'[j] = b[j] + sum_k W[j,k] shift[k]          (W as packed: bf16)
__global__ void __launch_bounds__(256) chain_fold_kernel(const ChainWeights w, const float* __restrict__ mod,
                                                         float* __restrict__ fold) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  constexpr long long nq = static_cast<long long>(kChainBlocks) * kChainQKVG, n13 = static_cast<long long>(kChainBlocks) * kChainW13;
  if (row >= nq + n13 + 64) return;
  const bf16* wr;
  const float *scale, *shift;
  float bias;
  float *ocs, *ob;
  if (row < nq) {
    const int blk = static_cast<int>(row / kChainQKVG);
    wr = w.wqkvg + row * kChainD;
    shift = mod + blk * 6 * kChainD;            // shift_msa
    scale = shift + kChainD;                    // scale_msa
    bias = w.bqkvg[row];
    ocs = fold + kFoldCsQ + row;
    ob = fold + kFoldBq + row;
  } else if (row < nq + n13) {
    const long long r = row - nq;
    const int blk = static_cast<int>(r / kChainW13);
    wr = w.w13 + r * kChainD;
    shift = mod + blk * 6 * kChainD + 3 * kChainD;  // shift_mlp
    scale = shift + kChainD;                        // scale_mlp
    bias = w.b13[r];
    ocs = fold + kFoldCs13 + r;
    ob = fold + kFoldB13 + r;
  } else {
    const long long r = row - nq - n13;
    wr = w.wvel + r * kChainD;
    scale = mod + kChainBlocks * 6 * kChainD;  // final adaLN: [scale, shift] (dit.py:37)
    shift = scale + kChainD;
    bias = w.bvel[r];
    ocs = fold + kFoldCsV + r;
    ob = fold + kFoldBv + r;
  }
  float a = 0.f, b = 0.f;
  const uint32_t* w2 = reinterpret_cast<const uint32_t*>(wr);
  for (int i = lane; i < kChainD / 2; i += 32) {
    const float2 wv = op16_unpack2(w2[i]);
    const float2 sc = *reinterpret_cast<const float2*>(scale + 2 * i);
    const float2 sh = *reinterpret_cast<const float2*>(shift + 2 * i);
    a = fmaf(wv.x, 1.0f + sc.x, fmaf(wv.y, 1.0f + sc.y, a));
    b = fmaf(wv.x, sh.x, fmaf(wv.y, sh.y, b));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane == 0) {
    *ocs = a;
    *ob = b + bias;
  }
}

// x [M][960] fp32 -> xb = bf16(x (1 + scale)), stats[M][30][2]; one warp per row, 8 lanes per 32-column chunk
__global__ void __launch_bounds__(256) chain_stats_cast_kernel(const float* __restrict__ x, int M,
                                                               const float* __restrict__ scale, bf16* __restrict__ xb,
                                                               float* __restrict__ stats) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * kChainD);
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int idx = it * 32 + lane;  // float4 index, 240 per row
    const bool live = idx < kChainD / 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), s = v;
    if (live) {
      v = xr[idx];
      s = __ldg(reinterpret_cast<const float4*>(scale) + idx);
    }
    float s1 = (v.x + v.y) + (v.z + v.w);
    float s2 = fmaf(v.x, v.x, v.y * v.y) + fmaf(v.z, v.z, v.w * v.w);
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (live) {
      uint2 pk;
      pk.x = bf2(v.x * (1.0f + s.x), v.y * (1.0f + s.y));
      pk.y = bf2(v.z * (1.0f + s.z), v.w * (1.0f + s.w));
      reinterpret_cast<uint2*>(xb + static_cast<long long>(row) * kChainD)[idx] = pk;
      if ((lane & 7) == 0) {
        *reinterpret_cast<float2*>(stats + (static_cast<long long>(row) * kChainParts + (idx >> 3)) * 2) = make_float2(s1, s2);
      }
    }
  }
}

__global__ void pack_rows_headpad_kernel(const float* __restrict__ src, int rows, int cols, int row_off,
                                         bf16* __restrict__ dst, int ld_dst) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(rows) * cols) return;
  const int r = static_cast<int>(i / cols), cc = static_cast<int>(i % cols);
  const int dr = (r / kChainHD) * kChainHDP + r % kChainHD + row_off;
  dst[static_cast<long long>(dr) * ld_dst + cc] = op16_from_float(src[i]);
}
__global__ void pack_vec_headpad_kernel(const float* __restrict__ src, int n, int off, float* __restrict__ dst) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[(i / kChainHD) * kChainHDP + i % kChainHD + off] = src[i];
}

bool make_map(CUtensorMap* out, const void* ptr, uint64_t cols, uint64_t rows, uint32_t box_rows) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t str[1] = {cols * 2};
  const uint32_t box[2] = {BK, box_rows};
  return tmap_tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, ptr, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace

cudaError_t launch_dit_chain(cudaStream_t st, const ChainWeights& w, const ChainBuffers& b, const ChainCall& c) {
  if (c.M < 1 || c.T < 1 || c.n_phases < 1 || c.n_phases > kChainMaxPhases || !c.mod || !c.fold || !b.ready) {
    return cudaErrorInvalidValue;
  }
  bool has_attn = false;
  for (int p = 0; p < c.n_phases; ++p) {
    if (c.kind[p] < CHAIN_QKVG || c.kind[p] > CHAIN_ATTN || c.blk[p] < 0 || c.blk[p] >= kChainBlocks) return cudaErrorInvalidValue;
    has_attn = has_attn || c.kind[p] == CHAIN_ATTN;
  }
  const int m_tiles = (c.M + BM - 1) / BM;
  if (has_attn) {
    const ChainAttn& a = c.attn;
    if (a.B < 1 || a.B * c.T != c.M || a.R < 1 || a.P < 1 || !a.ref_len || !a.ph_len || !a.kv_ref || !a.kv_text || !c.frames ||
        m_tiles > kChainMaxRowBlocks || !chain_attn_fits(c.T, a.R, a.P)) {
      return cudaErrorInvalidValue;
    }
  }
  static PerDeviceOnce attr_set;
  {
    const cudaError_t err = attr_set.run([] {
      const cudaError_t e0 = cudaFuncSetAttribute(dit_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
      return e0 != cudaSuccess ? e0 : cudaFuncSetAttribute(dit_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    });
    if (err != cudaSuccess) return err;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  ChainMaps maps;
  const uint64_t M = static_cast<uint64_t>(c.M);
  bool ok = make_map(&maps.a[0], b.xb, kChainD, M, BM) && make_map(&maps.a[1], b.ob, kChainH * kChainHDP, M, BM) &&
            make_map(&maps.a[2], b.hb, kChainFF, M, BM) &&
            make_map(&maps.w[CHAIN_QKVG], w.wqkvg, kChainD, static_cast<uint64_t>(kChainBlocks) * kChainQKVG, 64) &&
            make_map(&maps.w[CHAIN_OUT], w.wo, kChainH * kChainHDP, static_cast<uint64_t>(kChainBlocks) * kChainD, 64) &&
            make_map(&maps.w[CHAIN_W13], w.w13, kChainD, static_cast<uint64_t>(kChainBlocks) * kChainW13, 64) &&
            make_map(&maps.w[CHAIN_W2], w.w2, kChainFF, static_cast<uint64_t>(kChainBlocks) * kChainD, 64) &&
            make_map(&maps.w[CHAIN_VEL], w.wvel, kChainD, 64, 64);
  if (!ok) return cudaErrorInvalidValue;
  ChainDev d;
  d.w = w;
  d.b = b;
  d.c = c;
  d.m_tiles = m_tiles;
  d.q_tiles = (c.T + chain_attn::kRows - 1) / chain_attn::kRows;
  d.attn_items = 0;
  for (int m = 0; m < kChainMaxRowBlocks; ++m) d.attn_target[m] = 0;
  if (has_attn) {
    const ChainAttn& a = c.attn;
    auto make3 = [&](CUtensorMap* mp, const bf16* ptr, uint64_t rows, uint32_t box_rows) {
      const uint64_t dims[3] = {kChainHDP, kChainH, rows};
      const uint64_t str[2] = {kChainHDP * 2, kChainH * kChainHDP * 2};
      const uint32_t box[3] = {64, 1, box_rows};
      return tmap_tiled(mp, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, ptr, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    };
    const size_t half = static_cast<size_t>(3) * M * kChainH * kChainHDP;  // elements of one q|k|v buffer
    ok = make3(&maps.at.q, b.qkv, (c.qkv_db ? 6 : 3) * M, chain_attn::kRows) && make3(&maps.at.kv_self[0], b.qkv, 3 * M, 16) &&
         make3(&maps.at.kv_self[1], b.qkv + (c.qkv_db ? half : 0), 3 * M, 16) &&
         make3(&maps.at.ref, a.kv_ref, static_cast<uint64_t>(kChainBlocks) * 2 * a.B * a.R, 16) &&
         make3(&maps.at.text, a.kv_text, static_cast<uint64_t>(kChainBlocks) * 2 * a.B * a.P, 16);
    if (!ok) return cudaErrorInvalidValue;
    d.attn_items = a.B * d.q_tiles * kChainH;
    for (int u = 0; u < a.B; ++u) {
      for (int qt = 0; qt < d.q_tiles; ++qt) {
        const int r0 = u * c.T + qt * chain_attn::kRows;
        const int r1 = u * c.T + (qt * chain_attn::kRows + chain_attn::kRows < c.T ? qt * chain_attn::kRows + chain_attn::kRows : c.T) - 1;
        for (int m = r0 / BM; m <= r1 / BM; ++m) d.attn_target[m] += kChainH;
      }
    }
  } else {
    maps.at = chain_attn::Maps{};
  }
  // every CTA must be resident at the same time (the phases wait on each other): one CTA per SM, never more
  int most = 0;
  for (int p = 0; p < c.n_phases; ++p) {
    const int nt = c.kind[p] == CHAIN_QKVG ? 29 : (c.kind[p] == CHAIN_W13 ? 25 : (c.kind[p] == CHAIN_VEL ? 1 : 15));
    const int items = c.kind[p] == CHAIN_ATTN ? d.attn_items : nt * d.m_tiles;
    most = items > most ? items : most;
  }
  const int grid = most < num_sms ? most : num_sms;
  const cudaError_t le = has_attn ? launch_k(dit_chain_kernel<true>, dim3(grid), dim3(kThreads), kSmem, st, maps, d)
                                  : launch_k(dit_chain_kernel<false>, dim3(grid), dim3(kThreads), kSmem, st, maps, d);
  count_launch();
  return le != cudaSuccess ? le : cudaGetLastError();
}

cudaError_t chain_fold_table(cudaStream_t st, const ChainWeights& w, const float* mod, float* fold) {
  const long long rows = static_cast<long long>(kChainBlocks) * (kChainQKVG + kChainW13) + 64;
  const cudaError_t le = launch_k(chain_fold_kernel, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0, st, w, mod, fold);
  count_launch();
  return le != cudaSuccess ? le : cudaGetLastError();
}

cudaError_t chain_stats_cast(cudaStream_t st, const float* x, int M, const float* scale, bf16* xb, float* stats) {
  const cudaError_t le = launch_k(chain_stats_cast_kernel, dim3((M + 7) / 8), dim3(256), 0, st, x, M, scale, xb, stats);
  count_launch();
  return le != cudaSuccess ? le : cudaGetLastError();
}

cudaError_t pack_rows_headpad(cudaStream_t st, const float* src, int rows, int cols, int row_off, bf16* dst, int ld_dst) {
  const long long n = static_cast<long long>(rows) * cols;
  const cudaError_t le = launch_k(pack_rows_headpad_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, st, src,
                                  rows, cols, row_off, dst, ld_dst);
  count_launch();
  return le != cudaSuccess ? le : cudaGetLastError();
}
cudaError_t pack_vec_headpad(cudaStream_t st, const float* src, int n, int off, float* dst) {
  const cudaError_t le = launch_k(pack_vec_headpad_kernel, dim3((n + 255) / 256), dim3(256), 0, st, src, n, off, dst);
  count_launch();
  return le != cudaSuccess ? le : cudaGetLastError();
}

}  // namespace stts
