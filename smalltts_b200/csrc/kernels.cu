// Non-GEMM kernels of the engine.  See kernels.cuh for the contracts and reference citations.
#include "kernels.cuh"

#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include "launch.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace stts {

namespace {

thread_local cudaError_t last_launch_status = cudaSuccess;
#define STTS_LAUNCH_OK()                                                         \
  do {                                                                           \
    count_launch();                                                              \
    const cudaError_t _e = last_launch_status;                                   \
    return _e != cudaSuccess ? _e : cudaGetLastError();                          \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint2 pack_bf16x4(float a, float b, float c, float d) {
  uint2 r;
  r.x = op16_pack2(a, b);
  r.y = op16_pack2(c, d);
  return r;
}

// ------------------------------------------------------------------ row norms (warp per row, dim <= 1024)
template <bool kLayerNorm>
__global__ void __launch_bounds__(256) row_norm_kernel(const float* __restrict__ x, int rows, int rpb, int dim,
                                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                                       int ld_mod, float eps, bf16* __restrict__ out) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const int nv = dim >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * dim);
  // scale / shift do not depend on the row statistics: request them together with the row (one latency, not three)
  const int bq = row / rpb;
  const float4* scq = reinterpret_cast<const float4*>(scale + static_cast<long long>(kLayerNorm ? bq : 0) * ld_mod);
  const float4* shq = kLayerNorm ? reinterpret_cast<const float4*>(shift + static_cast<long long>(bq) * ld_mod) : nullptr;
  float4 v[8], wq[8], hq[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = lane + 32 * i;
    wq[i] = idx < nv ? __ldg(scq + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    hq[i] = (kLayerNorm && idx < nv) ? __ldg(shq + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = lane + 32 * i;
    v[i] = idx < nv ? xr[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
    s += kLayerNorm ? (v[i].x + v[i].y + v[i].z + v[i].w) : (v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w);
  }
  s = warp_sum(s);
  float mean = 0.f, rstd;
  if (kLayerNorm) {
    mean = s / dim;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (lane + 32 * i < nv) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += a * a + b * b + c * c + d * d;
      }
    }
    q = warp_sum(q);
    rstd = 1.0f / sqrtf(q / dim + eps);
  } else {
    rstd = 1.0f / sqrtf(s / dim + eps);
  }
  uint2* o = reinterpret_cast<uint2*>(out + static_cast<long long>(row) * dim);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nv) {
      const float4 w = wq[i];
      float4 r;
      if (kLayerNorm) {
        const float4 h = hq[i];
        r.x = (v[i].x - mean) * rstd * (1.f + w.x) + h.x;
        r.y = (v[i].y - mean) * rstd * (1.f + w.y) + h.y;
        r.z = (v[i].z - mean) * rstd * (1.f + w.z) + h.z;
        r.w = (v[i].w - mean) * rstd * (1.f + w.w) + h.w;
      } else {
        r.x = v[i].x * rstd * w.x;
        r.y = v[i].y * rstd * w.y;
        r.z = v[i].z * rstd * w.z;
        r.w = v[i].w * rstd * w.w;
      }
      o[idx] = pack_bf16x4(r.x, r.y, r.z, r.w);
    }
  }
}

// ------------------------------------------------------------------ head split + per-head RMSNorm + RoPE
struct HeadJobs {  // up to three head-split jobs in one launch (q, k, v of one projection): blockIdx.y selects
  int src_off[3];
  const float* norm_w[3];
  int rot[3];
  bf16* out[3];
};

template <int EPL>  // elements per lane: hd_pad / 32; one warp per (row, head)
__device__ __forceinline__ void head_split_body(const float* __restrict__ in, int ld_in, int rows, int rpb, int heads,
                                                int hd, float eps, const float* __restrict__ cos_t,
                                                const float* __restrict__ sin_t, int src_off, int rot,
                                                const float* __restrict__ norm_w, bf16* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long item = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (item >= static_cast<long long>(rows) * heads) return;
  const int row = static_cast<int>(item / heads);
  const int h = static_cast<int>(item % heads);
  const int d0 = lane * EPL;
  const float* src = in + static_cast<long long>(row) * ld_in + src_off + h * hd;
  // every load of the item is requested up front (values, norm weights, rotary table): one memory latency instead of
  // three dependent ones.  hd, src_off and ld_in are multiples of EPL (checked by the launcher), so lanes are either
  // fully inside or fully outside the head.
  const bool live = d0 < hd;
  float v[EPL], nw[EPL], cs[EPL / 2], sn[EPL / 2];
  const bool rotate = rot > 0 && d0 < rot;
  if (EPL == 4) {
    const float4 t = live ? *reinterpret_cast<const float4*>(src + d0) : make_float4(0.f, 0.f, 0.f, 0.f);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    if (norm_w != nullptr) {
      const float4 w = live ? __ldg(reinterpret_cast<const float4*>(norm_w + h * hd + d0)) : make_float4(0.f, 0.f, 0.f, 0.f);
      nw[0] = w.x; nw[1] = w.y; nw[2] = w.z; nw[3] = w.w;
    }
  } else {
    const float2 t = live ? *reinterpret_cast<const float2*>(src + d0) : make_float2(0.f, 0.f);
    v[0] = t.x; v[1] = t.y;
    if (norm_w != nullptr) {
      const float2 w = live ? __ldg(reinterpret_cast<const float2*>(norm_w + h * hd + d0)) : make_float2(0.f, 0.f);
      nw[0] = w.x; nw[1] = w.y;
    }
  }
  if (rotate) {
    const int pos = row % rpb;
#pragma unroll
    for (int e = 0; e < EPL / 2; ++e) {
      cs[e] = __ldg(cos_t + pos * (rot >> 1) + (d0 >> 1) + e);
      sn[e] = __ldg(sin_t + pos * (rot >> 1) + (d0 >> 1) + e);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < EPL; ++e) s += v[e] * v[e];
  if (norm_w != nullptr) {
    s = warp_sum(s);
    const float r = 1.0f / sqrtf(s / hd + eps);
#pragma unroll
    for (int e = 0; e < EPL; ++e) v[e] = v[e] * r * nw[e];
  }
  if (rotate) {
#pragma unroll
    for (int e = 0; e < EPL; e += 2) {
      const float c = cs[e >> 1], sn_ = sn[e >> 1];
      const float x0 = v[e], x1 = v[e + 1];
      v[e] = x0 * c - x1 * sn_;
      v[e + 1] = x1 * c + x0 * sn_;
    }
  }
  bf16* o = out + (static_cast<long long>(row) * heads + h) * (EPL * 32) + d0;
#pragma unroll
  for (int e = 0; e < EPL; e += 2) {
    *reinterpret_cast<uint32_t*>(o + e) = op16_pack2(v[e], v[e + 1]);
  }
}

template <int EPL>
__global__ void __launch_bounds__(256) head_split_kernel(const float* __restrict__ in, int ld_in, int rows, int rpb,
                                                         int heads, int hd, float eps, const float* __restrict__ cos_t,
                                                         const float* __restrict__ sin_t, const HeadJobs jobs) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const int job = blockIdx.y;
  const int src_off = job == 0 ? jobs.src_off[0] : (job == 1 ? jobs.src_off[1] : jobs.src_off[2]);
  const int rot = job == 0 ? jobs.rot[0] : (job == 1 ? jobs.rot[1] : jobs.rot[2]);
  const float* __restrict__ norm_w = job == 0 ? jobs.norm_w[0] : (job == 1 ? jobs.norm_w[1] : jobs.norm_w[2]);
  bf16* __restrict__ out = job == 0 ? jobs.out[0] : (job == 1 ? jobs.out[1] : jobs.out[2]);
  head_split_body<EPL>(in, ld_in, rows, rpb, heads, hd, eps, cos_t, sin_t, src_off, rot, norm_w, out);
}

// Cross K/V caches of all layers in one launch (dit.py:80-93): blockIdx.y = 2*layer + {0: k (k_norm_cross), 1: v};
// source columns y*width of the fused projection, destination cache + y*out_stride.
template <int EPL>
__global__ void __launch_bounds__(256) kv_split_kernel(const float* __restrict__ in, int ld_in, int rows, int heads,
                                                       int hd, int width, float eps, const float* __restrict__ knorm_all,
                                                       bf16* __restrict__ cache, long long out_stride) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const int y = blockIdx.y;
  const float* norm_w = (y & 1) ? nullptr : knorm_all + static_cast<long long>(y >> 1) * heads * hd;
  head_split_body<EPL>(in, ld_in, rows, 1, heads, hd, eps, nullptr, nullptr, y * width, 0, norm_w, cache + y * out_stride);
}

// ------------------------------------------------------------------ attention (warp MMA, keys split over warps)
// CTA = 48 query rows of one (batch, head): 3 row groups x 4 key splits = 12 warps.  The joint key sequence
// [self | ref | text] is laid out virtually with every segment padded to a multiple of 16 keys; a chunk is 64 virtual
// keys = four TMA boxes of 16 rows per operand half, so no thread ever computes a K/V address.  Warp (rg, split)
// scores its 16 rows against chunks split, split+4, ... with an online softmax and the four partial results of a row
// group are merged through shared memory.  At the DiT's 210 keys every warp sees exactly one chunk: the kernel is one
// TMA round trip + 128 warp MMAs + the merge deep, instead of a serial walk over all keys by two warps.
constexpr int kAttRG = 3, kAttSplits = 4;
constexpr int kAttRows = kAttRG * 16;
constexpr int kAttThreads = kAttRG * kAttSplits * 32;
constexpr int kAttChunk = 64;

struct AttnMaps {
  CUtensorMap q;     // [hd_pad, H, B*tq]      box [64, 1, 48]
  CUtensorMap k[3];  // [hd_pad, H, B*n_max]   box [64, 1, 16]
  CUtensorMap v[3];
};
struct AttnLens {
  const int* len[3];
  int n_max[3];
};

// byte offset of (row, dim) in a [HD/64 halves][ROWS rows][64] bf16 tile written by TMA with 128-byte swizzle
template <int ROWS>
__device__ __forceinline__ uint32_t att_sw(int row, int dim) {
  return static_cast<uint32_t>((dim >> 6) * (ROWS * 128) + row * 128 + ((((dim & 63) >> 3) ^ (row & 7)) << 4));
}

template <int HD>
__global__ void __launch_bounds__(kAttThreads, 1)
attention_kernel(const __grid_constant__ AttnMaps maps, const AttnLens lens, int tq, int H, int hd, int nseg,
                 const float* __restrict__ gate, int ld_gate, int gate_off, float scale_log2, bf16* __restrict__ out) {
  constexpr int NH = HD / 64;                         // 128-byte halves per row
  constexpr int Q_BYTES = NH * kAttRows * 128;        // 12 KB / 6 KB
  constexpr int OP_BYTES = NH * kAttChunk * 128;      // one K or V chunk: 16 KB / 8 KB
  constexpr int SPLIT_BYTES = 2 * OP_BYTES;           // K + V of one split (later: 3 x [16][HD] fp32 partial outputs)
  constexpr int O_BYTES = 16 * HD * 4;
  static_assert(kAttRG * O_BYTES <= SPLIT_BYTES, "merge buffers must fit in the K/V buffer of a split");
  extern __shared__ uint8_t smem_att_raw[];
  const uint32_t raw = ptx::smem_u32(smem_att_raw);
  uint8_t* smem = smem_att_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + ((Q_BYTES + 1023) & ~1023);
  float* ml = reinterpret_cast<float*>(sKV + kAttSplits * SPLIT_BYTES);  // [12 warps][16 rows][m, l]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ml + kAttRG * kAttSplits * 32);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;  // [4]

  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kAttRows;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = warp / kAttRG, rg = warp % kAttRG;

  if (threadIdx.x == 0) {
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < kAttSplits; ++i) ptx::mbar_init(&kv_full[i], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  ptx::pdl_wait();
  ptx::pdl_trigger();

  // virtual key layout: segment s occupies [es[s-1], es[s-1] + pad16(len_s))
  int len0 = lens.len[0] ? min(lens.len[0][b], lens.n_max[0]) : lens.n_max[0];
  int len1 = 0, len2 = 0;
  if (nseg > 1) len1 = lens.len[1] ? min(lens.len[1][b], lens.n_max[1]) : lens.n_max[1];
  if (nseg > 2) len2 = lens.len[2] ? min(lens.len[2][b], lens.n_max[2]) : lens.n_max[2];
  const int e0 = (len0 + 15) & ~15, e1 = e0 + ((len1 + 15) & ~15), e2 = e1 + ((len2 + 15) & ~15);
  const int n_chunks = (e2 + kAttChunk - 1) / kAttChunk;

  auto issue_chunk = [&](int chunk) {  // one thread: 4 groups x {K, V} x NH boxes of [64 dims x 16 rows]
    uint64_t* bar = &kv_full[split];
    ptx::mbar_expect_tx(bar, SPLIT_BYTES);
    uint8_t* dK = sKV + split * SPLIT_BYTES;
    uint8_t* dV = dK + OP_BYTES;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int vk = chunk * kAttChunk + g * 16;
      int s = 0, j = vk;
      if (vk >= e1) { s = 2; j = vk - e1; } else if (vk >= e0) { s = 1; j = vk - e0; }
      // groups past the end of the sequence are fetched from beyond the tensor: TMA fills them with zeros
      const int rowc = vk < e2 ? b * lens.n_max[s] + j : 0x40000000;
#pragma unroll
      for (int hf = 0; hf < NH; ++hf) {
        ptx::tma_load_3d(dK + hf * (kAttChunk * 128) + g * 2048, &maps.k[s], bar, hf * 64, h, rowc);
        ptx::tma_load_3d(dV + hf * (kAttChunk * 128) + g * 2048, &maps.v[s], bar, hf * 64, h, rowc);
      }
    }
  };

  if (warp == 0 && lane == 0) {
    ptx::mbar_expect_tx(q_full, Q_BYTES);
#pragma unroll
    for (int hf = 0; hf < NH; ++hf) ptx::tma_load_3d(sQ + hf * (kAttRows * 128), &maps.q, q_full, hf * 64, h, b * tq + q0);
  }
  if (rg == 0 && lane == 0 && split < n_chunks) issue_chunk(split);

  const bool active = q0 + rg * 16 < tq;  // row groups past the end of the sequence only keep the barriers company
  float o[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  const uint32_t qa = ptx::smem_u32(sQ);
  const uint32_t ka = ptx::smem_u32(sKV + split * SPLIT_BYTES);
  const uint32_t va = ka + OP_BYTES;

  ptx::mbar_wait(q_full, 0);  // every thread: the CTA never exits with the Q copy in flight
  int it = 0;
  for (int chunk = split; chunk < n_chunks; chunk += kAttSplits, ++it) {
    if (it > 0 && rg == 0 && lane == 0) issue_chunk(chunk);  // buffer released by the barrier that ended it-1
    ptx::mbar_wait(&kv_full[split], it & 1);
    if (active) {
      // S = Q K^T for 16 rows x 64 keys
      float sc[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
      const int id = lane >> 3, r8 = lane & 7;
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        uint32_t qf[4];
        ptx::ldmatrix_x4_addr(qf, qa + att_sw<kAttRows>(rg * 16 + (lane & 15), kk * 16 + (lane >> 4) * 8));
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {  // pairs of 8-key tiles
          uint32_t bfr[4];
          ptx::ldmatrix_x4_addr(bfr, ka + att_sw<kAttChunk>(jp * 16 + (id >> 1) * 8 + r8, kk * 16 + (id & 1) * 8));
          ptx::mma_16816(sc[2 * jp], qf, bfr[0], bfr[1]);
          ptx::mma_16816(sc[2 * jp + 1], qf, bfr[2], bfr[3]);
        }
      }
      // mask padding keys, online softmax (rows: lane/4 and lane/4 + 8; cols: 8j + (lane%4)*2 + {0,1})
      const int kbase = chunk * kAttChunk;
      auto valid = [&](int kv) { return kv < e0 ? kv < len0 : (kv < e1 ? kv - e0 < len1 : kv - e1 < len2); };
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kc = kbase + j * 8 + (lane & 3) * 2;
        if (!valid(kc)) sc[j][0] = sc[j][2] = -INFINITY;
        if (!valid(kc + 1)) sc[j][1] = sc[j][3] = -INFINITY;
        mx[0] = fmaxf(mx[0], fmaxf(sc[j][0], sc[j][1]));
        mx[1] = fmaxf(mx[1], fmaxf(sc[j][2], sc[j][3]));
      }
      float corr[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        const float m_new = fmaxf(m_run[r], mx[r]);  // finite: every chunk holds >= 1 valid key
        corr[r] = exp2f((m_run[r] - m_new) * scale_log2);
        m_run[r] = m_new;
        l_run[r] *= corr[r];
      }
      uint32_t pf[4][4];  // P as A fragments for the 4 k-steps of 16 keys
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float p0 = exp2f((sc[j][0] - m_run[0]) * scale_log2);
        const float p1 = exp2f((sc[j][1] - m_run[0]) * scale_log2);
        const float p2 = exp2f((sc[j][2] - m_run[1]) * scale_log2);
        const float p3 = exp2f((sc[j][3] - m_run[1]) * scale_log2);
        l_run[0] += p0 + p1;
        l_run[1] += p2 + p3;
        pf[j >> 1][(j & 1) * 2 + 0] = op16_pack2(p0, p1);
        pf[j >> 1][(j & 1) * 2 + 1] = op16_pack2(p2, p3);
      }
      if (it > 0) {
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) {
          o[i][0] *= corr[0]; o[i][1] *= corr[0];
          o[i][2] *= corr[1]; o[i][3] *= corr[1];
        }
      }
      // O += P V
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int np = 0; np < HD / 16; ++np) {  // pairs of 8-dim tiles
          uint32_t bfr[4];
          ptx::ldmatrix_x4_trans_addr(bfr, va + att_sw<kAttChunk>(kk * 16 + (id & 1) * 8 + r8, (np * 2 + (id >> 1)) * 8));
          ptx::mma_16816(o[2 * np], pf[kk], bfr[0], bfr[1]);
          ptx::mma_16816(o[2 * np + 1], pf[kk], bfr[2], bfr[3]);
        }
      }
    }
    // the three warps of this split are done with the K/V buffer (next chunk / merge buffers may overwrite it)
    ptx::named_bar_sync(1 + split, kAttRG * 32);
  }

  // ---- merge the four key splits of every row group through shared memory
  if (active) {
    float* ob = reinterpret_cast<float*>(sKV + split * SPLIT_BYTES + rg * O_BYTES);  // [16][HD], 8-float XOR swizzle
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int rl = (lane >> 2) + r * 8;
      float l = l_run[r];
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      if ((lane & 3) == 0) {
        ml[(warp * 16 + rl) * 2 + 0] = m_run[r];
        ml[(warp * 16 + rl) * 2 + 1] = l;
      }
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) {
        const int d = i * 8 + (lane & 3) * 2;
        *reinterpret_cast<float2*>(ob + rl * HD + (d ^ ((rl & 7) << 3))) = make_float2(o[i][2 * r], o[i][2 * r + 1]);
      }
    }
  }
  __syncthreads();
  for (int item = threadIdx.x; item < kAttRows * 8; item += kAttThreads) {
    const int row = item >> 3, c8 = item & 7;
    if (q0 + row >= tq) continue;
    const int g = row >> 4, rl = row & 15;
    float ms[kAttSplits], w[kAttSplits];
    float m_max = -INFINITY;
#pragma unroll
    for (int s = 0; s < kAttSplits; ++s) {
      ms[s] = ml[((s * kAttRG + g) * 16 + rl) * 2];
      m_max = fmaxf(m_max, ms[s]);
    }
    float L = 0.f;
#pragma unroll
    for (int s = 0; s < kAttSplits; ++s) {
      w[s] = m_max == -INFINITY ? 0.f : exp2f((ms[s] - m_max) * scale_log2);
      L += w[s] * ml[((s * kAttRG + g) * 16 + rl) * 2 + 1];
    }
    const float inv = L > 0.f ? 1.0f / L : 0.f;
    const long long m = static_cast<long long>(b) * tq + q0 + row;
    bf16* op = out + m * (H * HD) + h * HD;
    const float* gp = gate ? gate + m * ld_gate + gate_off + h * hd : nullptr;
#pragma unroll
    for (int k = 0; k < HD / 32; ++k) {
      const int d = 4 * c8 + 32 * k;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int s = 0; s < kAttSplits; ++s) {
        const float* ob = reinterpret_cast<const float*>(sKV + s * SPLIT_BYTES + g * O_BYTES);
        const float4 x = *reinterpret_cast<const float4*>(ob + rl * HD + (d ^ ((rl & 7) << 3)));
        acc.x = fmaf(w[s], x.x, acc.x); acc.y = fmaf(w[s], x.y, acc.y);
        acc.z = fmaf(w[s], x.z, acc.z); acc.w = fmaf(w[s], x.w, acc.w);
      }
      acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
      if (gp != nullptr) {
        if (d < hd) {  // hd is a multiple of 4
          const float4 gv = *reinterpret_cast<const float4*>(gp + d);
          acc.x = __fdividef(acc.x, 1.0f + __expf(-gv.x)); acc.y = __fdividef(acc.y, 1.0f + __expf(-gv.y));
          acc.z = __fdividef(acc.z, 1.0f + __expf(-gv.z)); acc.w = __fdividef(acc.w, 1.0f + __expf(-gv.w));
        } else {
          acc = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      *reinterpret_cast<uint2*>(op + d) = pack_bf16x4(acc.x, acc.y, acc.z, acc.w);
    }
  }
}

// ------------------------------------------------------------------ small dense layers
template <int MAXR>
__global__ void __launch_bounds__(256) gemv_rows_kernel(const float* __restrict__ x, int rows, int k,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        int n, int pre, int post, int chunk, unsigned tanh_chunks,
                                                        float* __restrict__ y, int ld_y) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int col = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (col >= n) return;
  float acc[MAXR];
#pragma unroll
  for (int r = 0; r < MAXR; ++r) acc[r] = 0.f;
  const float4* wr = reinterpret_cast<const float4*>(w + static_cast<long long>(col) * k);
  for (int i = lane; i < (k >> 2); i += 32) {
    const float4 wv = wr[i];
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
      if (r < rows) {
        float4 xv = reinterpret_cast<const float4*>(x + static_cast<long long>(r) * k)[i];
        if (pre == 1) {
          xv.x = xv.x / (1.f + expf(-xv.x)); xv.y = xv.y / (1.f + expf(-xv.y));
          xv.z = xv.z / (1.f + expf(-xv.z)); xv.w = xv.w / (1.f + expf(-xv.w));
        }
        acc[r] += wv.x * xv.x + wv.y * xv.y + wv.z * xv.z + wv.w * xv.w;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < MAXR; ++r) {
    if (r < rows) {
      float v = warp_sum(acc[r]);
      if (lane == 0) {
        if (bias != nullptr) v += bias[col];
        if (post == 1) v = v / (1.f + expf(-v));
        if (chunk > 0 && ((tanh_chunks >> (col / chunk)) & 1u)) v = tanhf(v);
        y[static_cast<long long>(r) * ld_y + col] = v;
      }
    }
  }
}

__global__ void time_features_kernel(const float* __restrict__ t, int rows, float* __restrict__ out) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const int r = blockIdx.x, j = threadIdx.x;  // 128 threads
  if (r >= rows) return;
  const float f = expf(static_cast<float>(j) * -0.07252236513367074f);  // ln(1e4)/127
  const float a = 1e3f * t[r] * f;
  out[r * 256 + j] = sinf(a);
  out[r * 256 + 128 + j] = cosf(a);
}

__global__ void embed_gather_kernel(const long long* __restrict__ ids, int rows, const float* __restrict__ table,
                                    int vocab, int dim, float* __restrict__ out) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const int r = blockIdx.x;
  long long id = ids[r];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const float4* src = reinterpret_cast<const float4*>(table + id * dim);
  float4* dst = reinterpret_cast<float4*>(out + static_cast<long long>(r) * dim);
  for (int i = threadIdx.x; i < (dim >> 2); i += blockDim.x) dst[i] = src[i];
}

__global__ void cast_bf16_kernel(const float* __restrict__ in, long long n, bf16* __restrict__ out) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(in + i);
    *reinterpret_cast<uint2*>(out + i) = pack_bf16x4(v.x, v.y, v.z, v.w);
  } else {
    for (long long j = i; j < n; ++j) out[j] = op16_from_float(in[j]);
  }
}

// x [n] fp32 -> bf16 [reps*n]: the CFG branches of the teacher share x_t (distill.py:75)
__global__ void repeat_cast_bf16_kernel(const float* __restrict__ in, long long n, int reps, bf16* __restrict__ out) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bf16 v = op16_from_float(in[i]);
  for (int r = 0; r < reps; ++r) out[r * n + i] = v;
}

// 3-way CFG combine (distill.py:97-103) + one deterministic DDIM step in v-parameterisation (train/utils.py:54-67):
//   v = v_c + s_text (v_c - v_no_text) + s_spk (v_c - v_no_spk);  x <- ca * x + cb * v
__global__ void cfg_ddim_update_kernel(float* __restrict__ x, const float* __restrict__ v3, long long n, float s_text,
                                       float s_spk, float ca, float cb) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float vc = v3[i], vt = v3[n + i], vs = v3[2 * n + i];
  const float v = vc + s_text * (vc - vt) + s_spk * (vc - vs);
  x[i] = ca * x[i] + cb * v;
}

__global__ void noise_mix_kernel(const float* __restrict__ xp, const float* __restrict__ nz, float alpha, float sigma,
                                 long long n, float* __restrict__ xt, bf16* __restrict__ xtb) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    const float v = alpha * xp[i] + sigma * nz[i];
    xt[i] = v;
    xtb[i] = op16_from_float(v);
  }
}

__global__ void dmd_update_kernel(const float* __restrict__ xt, const float* __restrict__ v, float alpha, float sigma,
                                  long long n, float* __restrict__ xp) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) xp[i] = alpha * xt[i] - sigma * v[i];
}

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__global__ void philox_normal_kernel(const unsigned long long* __restrict__ seed_ptr, unsigned long long stream_id, long long n,
                                     float* __restrict__ out) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i * 4 >= n) return;
  uint32_t c[4] = {static_cast<uint32_t>(i), static_cast<uint32_t>(i >> 32), static_cast<uint32_t>(stream_id),
                   static_cast<uint32_t>(stream_id >> 32)};
  const unsigned long long seed = *seed_ptr;  // device-resident so that a captured graph can be re-seeded
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  float z[4];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const float u1 = (static_cast<float>(c[2 * p]) + 1.0f) * 2.3283064365386963e-10f;  // (0, 1]
    const float u2 = static_cast<float>(c[2 * p + 1]) * 2.3283064365386963e-10f;
    const float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincosf(6.283185307179586f * u2, &sn, &cs);
    z[2 * p] = rad * cs;
    z[2 * p + 1] = rad * sn;
  }
  for (int j = 0; j < 4; ++j) {
    if (i * 4 + j < n) out[i * 4 + j] = z[j];
  }
}

// ------------------------------------------------------------------ vocoder ConvNeXt token mixer
// One CTA = TT output rows (+6 causal halo rows) x all C channels, staged in shared memory.
__global__ void __launch_bounds__(256) convnext_mix_kernel(const float* __restrict__ x, int T, int C, int TT,
                                                           const float* __restrict__ norm_w,
                                                           const float* __restrict__ conv_w,
                                                           const float* __restrict__ conv_b,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ ffn_norm_w, float eps,
                                                           float* __restrict__ y, bf16* __restrict__ a) {
  extern __shared__ __align__(16) float smx[];
  const int R = TT + 6;
  float* tile = smx;               // [R][C]
  float* inv1 = smx + R * C;       // [R]
  float* inv2 = inv1 + R;          // [TT]
  // per-channel parameters, staged BEFORE the dependency wait (weights do not depend on the predecessor): the conv
  // loop below walks C / 256 channels per thread, and a dependent global-load round trip per channel was most of this
  // kernel's time at C = 2048 (19 us for 5 MB of activations)
  float* cw = inv2 + ((TT + 3) & ~3);  // [C][7]
  float* cbs = cw + 7 * C;             // [C] conv bias
  float* gms = cbs + C;                // [C] layer scale
  float* nws = gms + C;                // [C] norm weight
  for (int i = threadIdx.x; i < 7 * C; i += 256) cw[i] = conv_w[i];
  for (int i = threadIdx.x; i < C; i += 256) {
    cbs[i] = conv_b[i];
    gms[i] = gamma[i];
    nws[i] = norm_w[i];
  }
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * TT;
  const int nrows = min(TT, T - t0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cv = C >> 2;                  // float4 per row
  const int lpr = cv < 32 ? cv : 32;      // lanes cooperating on one row (power of two)
  const int rpi = 32 / lpr;               // rows per warp iteration
  const int sub = lane / lpr, sl = lane % lpr;
  const float* xb = x + static_cast<long long>(b) * T * C;

  // phase 1: stage rows t0-6 .. t0+nrows-1 with cp.async (every 16-byte piece of the tile in flight at once: one
  // memory round trip instead of one per row), then per-row 1/rms from shared memory
  {
    const uint32_t dst0 = ptx::smem_u32(tile);
    for (int i = threadIdx.x; i < R * cv; i += 256) {
      const int r = i / cv, c4 = i - r * cv;
      const int t = t0 - 6 + r;
      const bool live = t >= 0 && r < nrows + 6;
      ptx::cp_async_16(dst0 + (r * C + c4 * 4) * 4, xb + static_cast<long long>(live ? t : 0) * C + c4 * 4, live ? 16u : 0u);
    }
    ptx::cp_async_commit();
    ptx::cp_async_wait<0>();
  }
  __syncthreads();
  for (int r0 = warp * rpi; r0 < R; r0 += 8 * rpi) {
    const int r = r0 + sub;
    const int t = t0 - 6 + r;
    const bool live = r < R && t >= 0 && r < nrows + 6;
    float s = 0.f;
    if (r < R) {
      const float4* src = reinterpret_cast<const float4*>(tile + r * C);
      for (int i = sl; i < cv; i += lpr) {
        const float4 v = src[i];
        s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
    }
    for (int o = lpr >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (r < R && sl == 0) inv1[r] = live ? 1.0f / sqrtf(s / C + eps) : 0.f;
  }
  __syncthreads();

  // phase 2: depthwise causal conv along time, one (channel, time segment) per thread
  const int nseg = C >= 256 ? 1 : 256 / C;
  const int seg_len = (nrows + nseg - 1) / nseg;
  const int seg = C >= 256 ? 0 : threadIdx.x / C;
  const int rs = 6 + seg * seg_len;                       // first output row (tile coordinates)
  const int re = min(6 + nrows, rs + seg_len);
  if (C >= 512) {
    // wide stages: one channel PAIR per thread and iteration, packed fp32 arithmetic (FFMA2; the scalar form below
    // contracts to the same fused multiply-adds, so the results are identical)
    for (int c = 2 * threadIdx.x; c < C; c += 512) {
      float2 w[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) w[j] = make_float2(cw[c * 7 + j], cw[(c + 1) * 7 + j]);
      const float2 cb = make_float2(cbs[c], cbs[c + 1]), gm = make_float2(gms[c], gms[c + 1]), nw = make_float2(nws[c], nws[c + 1]);
      float2 win[7];
      win[0] = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 1; j < 7; ++j) {
        const int r = rs - 7 + j;  // rows rs-6 .. rs-1
        win[j] = (r >= 0 && rs < re) ? ptx::f2_mul(ptx::f2_scale(*reinterpret_cast<const float2*>(tile + r * C + c), inv1[r]), nw)
                                     : make_float2(0.f, 0.f);
      }
      for (int r = rs; r < re; ++r) {
        float2* xp = reinterpret_cast<float2*>(tile + r * C + c);
        const float2 xv = *xp;
#pragma unroll
        for (int j = 0; j < 6; ++j) win[j] = win[j + 1];
        win[6] = ptx::f2_mul(ptx::f2_scale(xv, inv1[r]), nw);
        float2 acc = cb;
#pragma unroll
        for (int j = 0; j < 7; ++j) acc = ptx::f2_fma(w[j], win[j], acc);
        *xp = ptx::f2_fma(gm, acc, xv);
      }
    }
  } else
  for (int c = (C >= 256 ? threadIdx.x : threadIdx.x % C); c < C; c += 256) {
    float w[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) w[j] = cw[c * 7 + j];  // stride 7 words across lanes: conflict-free
    const float cb = cbs[c], gm = gms[c], nw = nws[c];
    float win[7];
    win[0] = 0.f;
#pragma unroll
    for (int j = 1; j < 7; ++j) {
      const int r = rs - 7 + j;  // rows rs-6 .. rs-1
      win[j] = (r >= 0 && rs < re) ? tile[r * C + c] * inv1[r] * nw : 0.f;
    }
    if (nseg > 1) __syncthreads();  // all warm-up reads precede in-place writes of neighbouring segments
    for (int r = rs; r < re; ++r) {
      const float xv = tile[r * C + c];
#pragma unroll
      for (int j = 0; j < 6; ++j) win[j] = win[j + 1];
      win[6] = xv * inv1[r] * nw;
      float acc = cb;
#pragma unroll
      for (int j = 0; j < 7; ++j) acc += w[j] * win[j];
      tile[r * C + c] = xv + gm * acc;
    }
  }
  __syncthreads();

  // phase 3: 1/rms of the updated rows
  for (int r0 = warp * rpi; r0 < nrows; r0 += 8 * rpi) {
    const int r = r0 + sub;
    float s = 0.f;
    if (r < nrows) {
      const float4* src = reinterpret_cast<const float4*>(tile + (r + 6) * C);
      for (int i = sl; i < cv; i += lpr) {
        const float4 v = src[i];
        s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
    }
    for (int o = lpr >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (r < nrows && sl == 0) inv2[r] = 1.0f / sqrtf(s / C + eps);
  }
  __syncthreads();

  // phase 4: coalesced stores of y (fp32) and a = rmsnorm(y) (bf16)
  const long long obase = (static_cast<long long>(b) * T + t0) * C;
  for (int i = threadIdx.x; i < nrows * cv; i += 256) {
    const int r = i / cv, c4 = i % cv;
    const float4 v = reinterpret_cast<const float4*>(tile + (r + 6) * C)[c4];
    reinterpret_cast<float4*>(y + obase)[i] = v;
    const float4 fw = reinterpret_cast<const float4*>(ffn_norm_w)[c4];
    const float s = inv2[r];
    reinterpret_cast<uint2*>(a + obase)[i] = pack_bf16x4(v.x * s * fw.x, v.y * s * fw.y, v.z * s * fw.z, v.w * s * fw.w);
  }
}

// Token mixer for C = 128 / 256 with per-thread row statistics (no shuffle chains; C+4 pitch keeps float4 row reads
// bank-conflict free) and cp.async staging.  Same contract as convnext_mix_kernel.
template <int C>
__global__ void __launch_bounds__(256) convnext_mix_rows_kernel(const float* __restrict__ x, int T,
                                                                const float* __restrict__ norm_w,
                                                                const float* __restrict__ conv_w,
                                                                const float* __restrict__ conv_b,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ ffn_norm_w, float eps,
                                                                float* __restrict__ y, bf16* __restrict__ a) {
  constexpr int TT = 16384 / C;  // 128 or 64 output rows per CTA
  constexpr int R = TT + 6, P = C + 4, CV = C / 4;
  constexpr int NSEG = 256 / C > 0 ? 256 / C : 1;  // time segments per channel (2 or 1)
  constexpr int SEGLEN = TT / NSEG;
  extern __shared__ __align__(16) float smr[];
  float* tile = smr;          // [R][P]
  float* inv1 = smr + R * P;  // [R]
  float* inv2 = inv1 + R;     // [TT]
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const int b = blockIdx.y, t0 = blockIdx.x * TT, tid = threadIdx.x;
  const int nrows = min(TT, T - t0);
  const float* xb = x + static_cast<long long>(b) * T * C;
  {
    const uint32_t dst0 = ptx::smem_u32(tile);
    for (int i = tid; i < R * CV; i += 256) {
      const int r = i / CV, c4 = i % CV;
      const int t = t0 - 6 + r;
      const bool ok = t >= 0 && r < nrows + 6;
      ptx::cp_async_16(dst0 + (r * P + c4 * 4) * 4, xb + static_cast<long long>(ok ? t : 0) * C + c4 * 4, ok ? 16u : 0u);
    }
    ptx::cp_async_commit();
    ptx::cp_async_wait<0>();
  }
  __syncthreads();
  if (tid < R) {
    const float4* row = reinterpret_cast<const float4*>(tile + tid * P);
    float2 s01 = make_float2(0.f, 0.f), s23 = make_float2(0.f, 0.f);  // packed fp32 (FFMA2): same sums, half the issue slots
#pragma unroll 8
    for (int j = 0; j < CV; ++j) {
      const float4 v = row[j];
      s01 = ptx::f2_fma(make_float2(v.x, v.y), make_float2(v.x, v.y), s01);
      s23 = ptx::f2_fma(make_float2(v.z, v.w), make_float2(v.z, v.w), s23);
    }
    const float s0 = s01.x, s1 = s01.y, s2 = s23.x, s3 = s23.y;
    inv1[tid] = rsqrtf((s0 + s1 + s2 + s3) * (1.0f / C) + eps);
  }
  __syncthreads();
  {
    // one (channel PAIR, time segment) per thread; packed fp32 arithmetic, bit-identical to the scalar form
    constexpr int NP = C / 2, NSEG2 = 256 / NP, SEGLEN2 = TT / NSEG2;
    const int c = 2 * (tid % NP), seg = tid / NP;
    const int rs = 6 + seg * SEGLEN2;
    float2 w[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) w[j] = make_float2(conv_w[c * 7 + j] * norm_w[c], conv_w[(c + 1) * 7 + j] * norm_w[c + 1]);
    const float2 cb = make_float2(conv_b[c], conv_b[c + 1]), gm = make_float2(gamma[c], gamma[c + 1]);
    float2 win[7];
    win[0] = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 1; j < 7; ++j) {
      win[j] = ptx::f2_scale(*reinterpret_cast<const float2*>(tile + (rs - 7 + j) * P + c), inv1[rs - 7 + j]);
    }
    __syncthreads();  // every warm-up read precedes the in-place writes of the previous segment
#pragma unroll 8
    for (int r = rs; r < rs + SEGLEN2; ++r) {
      float2* xp = reinterpret_cast<float2*>(tile + r * P + c);
      const float2 xv = *xp;
#pragma unroll
      for (int j = 0; j < 6; ++j) win[j] = win[j + 1];
      win[6] = ptx::f2_scale(xv, inv1[r]);
      float2 a0 = ptx::f2_fma(w[0], win[0], cb), a1 = ptx::f2_mul(w[1], win[1]);
      a0 = ptx::f2_fma(w[2], win[2], a0); a1 = ptx::f2_fma(w[3], win[3], a1);
      a0 = ptx::f2_fma(w[4], win[4], a0); a1 = ptx::f2_fma(w[5], win[5], a1);
      a0 = ptx::f2_fma(w[6], win[6], a0);
      *xp = ptx::f2_fma(gm, ptx::f2_add(a0, a1), xv);
    }
  }
  __syncthreads();
  if (tid < TT) {
    const float4* row = reinterpret_cast<const float4*>(tile + (tid + 6) * P);
    float2 s01 = make_float2(0.f, 0.f), s23 = make_float2(0.f, 0.f);  // packed fp32 (FFMA2): same sums, half the issue slots
#pragma unroll 8
    for (int j = 0; j < CV; ++j) {
      const float4 v = row[j];
      s01 = ptx::f2_fma(make_float2(v.x, v.y), make_float2(v.x, v.y), s01);
      s23 = ptx::f2_fma(make_float2(v.z, v.w), make_float2(v.z, v.w), s23);
    }
    const float s0 = s01.x, s1 = s01.y, s2 = s23.x, s3 = s23.y;
    inv2[tid] = rsqrtf((s0 + s1 + s2 + s3) * (1.0f / C) + eps);
  }
  __syncthreads();
  const long long obase = (static_cast<long long>(b) * T + t0) * C;
  for (int i = tid; i < nrows * CV; i += 256) {
    const int r = i / CV, c4 = i % CV;
    const float4 v = *reinterpret_cast<const float4*>(tile + (r + 6) * P + c4 * 4);
    reinterpret_cast<float4*>(y + obase)[i] = v;
    const float4 fw = reinterpret_cast<const float4*>(ffn_norm_w)[c4];
    const float s = inv2[r];
    const float2 q0 = ptx::f2_mul(ptx::f2_scale(make_float2(v.x, v.y), s), make_float2(fw.x, fw.y));
    const float2 q1 = ptx::f2_mul(ptx::f2_scale(make_float2(v.z, v.w), s), make_float2(fw.z, fw.w));
    reinterpret_cast<uint2*>(a + obase)[i] = pack_bf16x4(q0.x, q0.y, q1.x, q1.y);
  }
}

// 256 outputs per CTA; the 262 input rows are staged in shared memory with cp.async (row pitch C+4 words: 16-byte
// aligned rows whose float4 reads by consecutive threads are bank-conflict free).  One output per thread: 7 taps x C
// channels as float4 FMAs against weights read as warp-uniform (broadcast) float4s -- a quarter of the shared-memory
// instructions of the scalar version, which was LSU-bound at 124 us (HBM floor of the 246 MB read: 38 us).
__global__ void __launch_bounds__(256) head_conv_kernel(const float* __restrict__ x, int T, int C,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        float* __restrict__ out) {
  extern __shared__ __align__(16) float sh[];
  float* sw = sh;            // [7][C] tap-major
  float* sx = sh + 7 * C;    // [262][C + 4]
  const int P = C + 4;
  for (int i = threadIdx.x; i < 7 * C; i += 256) {  // weights: before the dependency wait
    const int j = i / C, c = i % C;
    sw[i] = w[c * 7 + j];
  }
  const float b0 = bias[0];
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * 256;
  const float* xb = x + static_cast<long long>(b) * T * C;
  const int cv = C >> 2;
  const uint32_t dst0 = ptx::smem_u32(sx);
  for (int i = threadIdx.x; i < 262 * cv; i += 256) {
    const int r = i / cv, c4 = i % cv;
    const int t = t0 - 6 + r;
    const bool live = t >= 0 && t < T;
    ptx::cp_async_16(dst0 + (r * P + 4 * c4) * 4, xb + static_cast<long long>(live ? t : 0) * C + 4 * c4, live ? 16u : 0u);
  }
  ptx::cp_async_commit();
  ptx::cp_async_wait<0>();
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t >= T) return;
  float2 a01 = make_float2(b0, 0.f), a23 = make_float2(0.f, 0.f);  // packed fp32: same four partial sums as before
  for (int j = 0; j < 7; ++j) {
    const float4* row = reinterpret_cast<const float4*>(sx + (threadIdx.x + j) * P);
    const float4* wr = reinterpret_cast<const float4*>(sw + j * C);
#pragma unroll 8
    for (int c = 0; c < cv; ++c) {
      const float4 v = row[c], k = wr[c];
      a01 = ptx::f2_fma(make_float2(v.x, v.y), make_float2(k.x, k.y), a01);
      a23 = ptx::f2_fma(make_float2(v.z, v.w), make_float2(k.z, k.w), a23);
    }
  }
  out[static_cast<long long>(b) * T + t] = (a01.x + a01.y) + (a23.x + a23.y);
}

// Codec encoder stem (hf:300-312): causal Conv1d(1 -> C, k=7) on raw audio.  One thread per sample; the thread
// writes its C outputs as whole 128-byte lines.
template <int C>
__global__ void __launch_bounds__(256) audio_stem_conv_kernel(const float* __restrict__ audio, int N,
                                                              const float* __restrict__ w, const float* __restrict__ bias,
                                                              float* __restrict__ out) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  __shared__ float sw[7 * C + C];
  for (int i = threadIdx.x; i < 7 * C; i += 256) sw[(i % 7) * C + i / 7] = w[i];  // tap-major
  for (int i = threadIdx.x; i < C; i += 256) sw[7 * C + i] = bias[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= N) return;
  const float* a = audio + static_cast<long long>(b) * N;
  float x[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) x[j] = (t - 6 + j >= 0) ? a[t - 6 + j] : 0.f;
  float4* o = reinterpret_cast<float4*>(out + (static_cast<long long>(b) * N + t) * C);
#pragma unroll
  for (int c4 = 0; c4 < C / 4; ++c4) {
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float acc = sw[7 * C + 4 * c4 + k];
#pragma unroll
      for (int j = 0; j < 7; ++j) acc = fmaf(sw[j * C + 4 * c4 + k], x[j], acc);
      v[k] = acc;
    }
    o[c4] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// Polyphase sinc resampler == torchaudio Resample as the reference configures it (infer/utils.py:7-23: kaiser window,
// lowpass_filter_width 1024, rolloff 0.94).  out[n*up + j] = sum_k bank[j][k] * xpad[n*down + k], xpad = x shifted by
// `width` zeros (torchaudio functional._apply_sinc_resample_kernel: F.pad + conv1d(stride = down)).  A CTA owns 32
// consecutive frames n (one per lane) and 8 phases j (one per warp): the (31*down + K)-sample input window is staged
// in shared memory once and read K times by every warp; the lanes of a warp read the same filter tap (one broadcast
// load) and inputs `down` apart.  Eight independent accumulators keep the fp32 summation error of the ~4k-tap
// filters near 1e-7 and give the FMA pipe independent work.
__global__ void __launch_bounds__(256) resample_kernel(const float* __restrict__ x, int N, const float* __restrict__ bank,
                                                       int K, int width, int down, int up, float* __restrict__ out,
                                                       int out_len) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  extern __shared__ float sx[];
  const int b = blockIdx.z;
  const int n0 = blockIdx.x * 32;
  const int span = 31 * down + K;
  const float* xb = x + static_cast<long long>(b) * N;
  const long long base = static_cast<long long>(n0) * down - width;
  for (int i = threadIdx.x; i < span; i += 256) {
    const long long g = base + i;
    sx[i] = (g >= 0 && g < N) ? xb[g] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.y * 8 + (threadIdx.x >> 5);
  const long long o = static_cast<long long>(n0 + lane) * up + j;
  if (j >= up || o >= out_len) return;
  const float* f = bank + static_cast<long long>(j) * K;
  const float* xs = sx + lane * down;
  float acc[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) acc[u] = 0.f;
  int k = 0;
  for (; k + 8 <= K; k += 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = fmaf(__ldg(f + k + u), xs[k + u], acc[u]);
  }
  for (; k < K; ++k) acc[0] = fmaf(__ldg(f + k), xs[k], acc[0]);
  out[static_cast<long long>(b) * out_len + o] =
      ((acc[0] + acc[4]) + (acc[2] + acc[6])) + ((acc[1] + acc[5]) + (acc[3] + acc[7]));
}

// ------------------------------------------------------------------ packing
// strided causal Conv1d weight [O, C, 2r] -> two-tap GEMM operand over rows regrouped r at a time:
// dst[o, tap*r*C + j*C + c] = w[o, c, tap*r + j]
__global__ void pack_conv_strided_kernel(const float* __restrict__ src, int O, int C, int r, bf16* __restrict__ dst) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long per_o = static_cast<long long>(2) * r * C;
  if (i >= O * per_o) return;
  const int o = static_cast<int>(i / per_o);
  const int rem = static_cast<int>(i % per_o);
  const int tap = rem / (r * C), j = (rem % (r * C)) / C, c = rem % C;
  dst[i] = op16_from_float(src[(static_cast<long long>(o) * C + c) * (2 * r) + tap * r + j]);
}

__device__ __forceinline__ int map_row(int r, int mode) {
  if (mode == ROW_INTERLEAVE16_LO) return (r >> 4) * 32 + (r & 15);
  if (mode == ROW_INTERLEAVE16_HI) return (r >> 4) * 32 + 16 + (r & 15);
  if (mode == ROW_GROUPPAD_60_64) return (r / 60) * 64 + r % 60;
  return r;
}

__global__ void pack_matrix_kernel(const float* __restrict__ src, int rows, int cols, float scale, int row_mode,
                                   int row_off, int col_mode, int col_off, bf16* __restrict__ dst, int ld_dst) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(rows) * cols) return;
  const int r = static_cast<int>(i / cols), c = static_cast<int>(i % cols);
  const int dr = map_row(r, row_mode) + row_off;
  const int dc = (col_mode == COL_HEADPAD_120_128 ? (c / 120) * 128 + c % 120 : c) + col_off;
  dst[static_cast<long long>(dr) * ld_dst + dc] = op16_from_float(scale * src[i]);
}

__global__ void pack_conv_taps_kernel(const float* __restrict__ src, int O, int cin, int taps, int kp, int opg,
                                      int group_pitch, bf16* __restrict__ dst, int ld_dst) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(O) * cin * taps) return;
  const int tap = static_cast<int>(i % taps);
  const int c = static_cast<int>((i / taps) % cin);
  const int o = static_cast<int>(i / (static_cast<long long>(taps) * cin));
  const int dr = (o / opg) * group_pitch + o % opg;
  dst[static_cast<long long>(dr) * ld_dst + tap * kp + c] = op16_from_float(src[i]);
}

__global__ void pack_conv_dense_tiles_kernel(const float* __restrict__ src, int taps, bf16* __restrict__ dst) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(960) * 60 * taps) return;
  const int tap = static_cast<int>(i % taps);
  const int c = static_cast<int>((i / taps) % 60);
  const int o = static_cast<int>(i / (static_cast<long long>(taps) * 60));
  const int kk = (o / 60 - o / 64) * 64 + c;
  dst[static_cast<long long>(o) * (taps * 128) + tap * 128 + kk] = op16_from_float(src[i]);
}

__global__ void pack_convtr_kernel(const float* __restrict__ src, int cin, int cout, int r, bf16* __restrict__ dst) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int kk = 2 * r;
  if (i >= static_cast<long long>(cin) * cout * kk) return;
  const int jj = static_cast<int>(i % kk);
  const int o = static_cast<int>((i / kk) % cout);
  const int c = static_cast<int>(i / (static_cast<long long>(kk) * cout));
  const int tap = jj / r, j = jj % r;
  dst[(static_cast<long long>(j) * cout + o) * (2 * cin) + tap * cin + c] = op16_from_float(src[i]);
}

__global__ void pack_vector_kernel(const float* __restrict__ src, int n, float scale, int row_mode, int off,
                                   float* __restrict__ dst) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[map_row(i, row_mode) + off] = scale * src[i];
}

__global__ void cast_f16_kernel(const float* __restrict__ src, long long n, float scale, __half* __restrict__ dst) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2half_rn(scale * src[i]);
}

__global__ void tile_vector_kernel(const float* __restrict__ src, int n, int reps, float* __restrict__ dst) {
  ptx::pdl_wait();
  ptx::pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * reps) dst[i] = src[i % n];
}

inline unsigned blocks_for(long long n, int per) { return static_cast<unsigned>((n + per - 1) / per); }

}  // namespace

// ==================================================================== launchers
cudaError_t ln_mod_bf16(cudaStream_t st, const float* x, int rows, int rpb, int dim, const float* scale,
                        const float* shift, int ld_mod, float eps, bf16* out) {
  if (dim % 4 != 0 || dim > 1024) return cudaErrorInvalidValue;
  last_launch_status = launch_k(row_norm_kernel<true>, dim3(blocks_for(rows, 8)), dim3(256), 0, st, x, rows, rpb, dim, scale, shift, ld_mod, eps, out);
  STTS_LAUNCH_OK();
}

cudaError_t rms_norm_bf16(cudaStream_t st, const float* x, int rows, int dim, const float* w, float eps, bf16* out) {
  if (dim % 4 != 0 || dim > 1024) return cudaErrorInvalidValue;
  last_launch_status = launch_k(row_norm_kernel<false>, dim3(blocks_for(rows, 8)), dim3(256), 0, st, x, rows, 1, dim, w, nullptr, 0, eps, out);
  STTS_LAUNCH_OK();
}

namespace {
cudaError_t head_split_launch(cudaStream_t st, const float* in, int ld_in, int rows, int rpb, int heads, int hd,
                              int hd_pad, float eps, const float* cos_t, const float* sin_t, const HeadJobs& jobs,
                              int njobs) {
  dim3 grid(blocks_for(static_cast<long long>(rows) * heads, 8), njobs);
  const int epl = hd_pad / 32;
  if ((hd % epl) != 0 || (ld_in % epl) != 0 || (reinterpret_cast<uintptr_t>(in) & 15) != 0) return cudaErrorInvalidValue;
  for (int j = 0; j < njobs; ++j) {
    if ((jobs.src_off[j] % epl) != 0 || (jobs.rot[j] % epl) != 0) return cudaErrorInvalidValue;
  }
  if (hd_pad == 128) {
    last_launch_status = launch_k(head_split_kernel<4>, grid, dim3(256), 0, st, in, ld_in, rows, rpb, heads, hd, eps, cos_t, sin_t, jobs);
  } else if (hd_pad == 64) {
    last_launch_status = launch_k(head_split_kernel<2>, grid, dim3(256), 0, st, in, ld_in, rows, rpb, heads, hd, eps, cos_t, sin_t, jobs);
  } else {
    return cudaErrorInvalidValue;
  }
  STTS_LAUNCH_OK();
}
}  // namespace

cudaError_t head_split_bf16(cudaStream_t st, const float* in, int ld_in, int src_off, int rows, int rpb, int heads,
                            int hd, int hd_pad, const float* norm_w, float eps, int rot, const float* cos_t,
                            const float* sin_t, bf16* out) {
  HeadJobs j = {};
  j.src_off[0] = src_off; j.norm_w[0] = norm_w; j.rot[0] = rot; j.out[0] = out;
  return head_split_launch(st, in, ld_in, rows, rpb, heads, hd, hd_pad, eps, cos_t, sin_t, j, 1);
}

cudaError_t kv_split_bf16(cudaStream_t st, const float* in, int ld_in, int rows, int layers, int heads, int hd,
                          int hd_pad, float eps, const float* knorm_all, bf16* cache, long long out_stride) {
  if (hd_pad != 128) return cudaErrorInvalidValue;
  dim3 grid(blocks_for(static_cast<long long>(rows) * heads, 8), 2 * layers);
  last_launch_status = launch_k(kv_split_kernel<4>, grid, dim3(256), 0, st, in, ld_in, rows, heads, hd, heads * hd, eps, knorm_all, cache, out_stride);
  STTS_LAUNCH_OK();
}

cudaError_t head_split_qkv_bf16(cudaStream_t st, const float* in, int ld_in, int stride_off, int rows, int rpb,
                                int heads, int hd, int hd_pad, const float* q_norm, const float* k_norm, float eps,
                                int rot, const float* cos_t, const float* sin_t, bf16* q, bf16* k, bf16* v) {
  HeadJobs j = {};
  j.src_off[0] = 0;              j.norm_w[0] = q_norm;  j.rot[0] = rot; j.out[0] = q;
  j.src_off[1] = stride_off;     j.norm_w[1] = k_norm;  j.rot[1] = rot; j.out[1] = k;
  j.src_off[2] = 2 * stride_off; j.norm_w[2] = nullptr; j.rot[2] = 0;   j.out[2] = v;
  return head_split_launch(st, in, ld_in, rows, rpb, heads, hd, hd_pad, eps, cos_t, sin_t, j, 3);
}

cudaError_t attention_bf16(cudaStream_t st, const bf16* q, int B, int tq, int H, int hd, int hd_pad, const AttnSeg* segs,
                           int nseg, const float* gate, int ld_gate, int gate_off, bf16* out) {
  if (nseg < 1 || nseg > 3 || (hd & 3) != 0 || (hd_pad != 64 && hd_pad != 128)) return cudaErrorInvalidValue;
  if (gate != nullptr && (((ld_gate | gate_off) & 3) != 0 || (reinterpret_cast<uintptr_t>(gate) & 15) != 0)) {
    return cudaErrorInvalidValue;
  }
  AttnMaps maps;
  AttnLens lens = {};
  const uint64_t row_bytes = static_cast<uint64_t>(H) * hd_pad * 2;
  auto make = [&](CUtensorMap* m, const bf16* p, long long rows, uint32_t box_rows) {
    const uint64_t dims[3] = {static_cast<uint64_t>(hd_pad), static_cast<uint64_t>(H), static_cast<uint64_t>(rows)};
    const uint64_t str[2] = {static_cast<uint64_t>(hd_pad) * 2, row_bytes};
    const uint32_t box[3] = {64, 1, box_rows};
    return tmap_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, p, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
  };
  if (!make(&maps.q, q, static_cast<long long>(B) * tq, kAttRows)) return cudaErrorInvalidValue;
  for (int i = 0; i < 3; ++i) {
    const AttnSeg& sg = segs[i < nseg ? i : 0];
    if (sg.k == nullptr || sg.v == nullptr || sg.n_max < 1) return cudaErrorInvalidValue;
    if (!make(&maps.k[i], sg.k, static_cast<long long>(B) * sg.n_max, 16)) return cudaErrorInvalidValue;
    if (!make(&maps.v[i], sg.v, static_cast<long long>(B) * sg.n_max, 16)) return cudaErrorInvalidValue;
    lens.len[i] = i < nseg ? sg.len : nullptr;
    lens.n_max[i] = i < nseg ? sg.n_max : 0;
  }
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(hd));
  dim3 grid((tq + kAttRows - 1) / kAttRows, H, B);
  auto smem_for = [](int hdp) {
    const int nh = hdp / 64;
    const int qb = (nh * kAttRows * 128 + 1023) & ~1023;
    return 1024 + qb + kAttSplits * 2 * nh * kAttChunk * 128 + kAttRG * kAttSplits * 32 * 4 + 64;
  };
  if (hd_pad == 128) {
    const int smem = smem_for(128);
    static PerDeviceOnce once;
    const cudaError_t e = once.run([smem] {
      return cudaFuncSetAttribute(attention_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    });
    if (e != cudaSuccess) return e;
    last_launch_status = launch_k(attention_kernel<128>, dim3(grid), dim3(kAttThreads), smem, st, maps, lens, tq, H, hd, nseg, gate, ld_gate, gate_off, scale_log2, out);
  } else {
    const int smem = smem_for(64);
    static PerDeviceOnce once;
    const cudaError_t e = once.run([smem] {
      return cudaFuncSetAttribute(attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    });
    if (e != cudaSuccess) return e;
    last_launch_status = launch_k(attention_kernel<64>, dim3(grid), dim3(kAttThreads), smem, st, maps, lens, tq, H, hd, nseg, gate, ld_gate, gate_off, scale_log2, out);
  }
  STTS_LAUNCH_OK();
}

cudaError_t gemv_rows(cudaStream_t st, const float* x, int rows, int k, const float* w, const float* b, int n, int pre,
                      int post, int chunk, unsigned tanh_chunks, float* y, int ld_y) {
  if (k % 4 != 0) return cudaErrorInvalidValue;
  for (int r0 = 0; r0 < rows; r0 += 8) {
    const int nr = rows - r0 < 8 ? rows - r0 : 8;
    last_launch_status = launch_k(gemv_rows_kernel<8>, dim3(blocks_for(n, 8)), dim3(256), 0, st, x + static_cast<long long>(r0) * k, nr, k, w, b, n, pre, post,
                                                         chunk, tanh_chunks, y + static_cast<long long>(r0) * ld_y, ld_y);
    count_launch();
  }
  return cudaGetLastError();
}

cudaError_t time_features(cudaStream_t st, const float* t, int rows, float* out) {
  last_launch_status = launch_k(time_features_kernel, dim3(rows), dim3(128), 0, st, t, rows, out);
  STTS_LAUNCH_OK();
}

cudaError_t embed_gather(cudaStream_t st, const long long* ids, int rows, const float* table, int vocab, int dim,
                         float* out) {
  last_launch_status = launch_k(embed_gather_kernel, dim3(rows), dim3(128), 0, st, ids, rows, table, vocab, dim, out);
  STTS_LAUNCH_OK();
}

cudaError_t cast_bf16(cudaStream_t st, const float* in, long long n, bf16* out) {
  last_launch_status = launch_k(cast_bf16_kernel, dim3(blocks_for(n, 1024)), dim3(256), 0, st, in, n, out);
  STTS_LAUNCH_OK();
}

cudaError_t repeat_cast_bf16(cudaStream_t st, const float* in, long long n, int reps, bf16* out) {
  last_launch_status = launch_k(repeat_cast_bf16_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, st, in, n, reps, out);
  STTS_LAUNCH_OK();
}

cudaError_t cfg_ddim_update(cudaStream_t st, float* x, const float* v3, long long n, float s_text, float s_spk, float ca,
                            float cb) {
  last_launch_status = launch_k(cfg_ddim_update_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, st, x, v3, n, s_text, s_spk, ca, cb);
  STTS_LAUNCH_OK();
}

cudaError_t noise_mix(cudaStream_t st, const float* x_pred, const float* noise, float alpha, float sigma, long long n,
                      float* x_t, bf16* x_t_bf16) {
  last_launch_status = launch_k(noise_mix_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, st, x_pred, noise, alpha, sigma, n, x_t, x_t_bf16);
  STTS_LAUNCH_OK();
}

cudaError_t dmd_update(cudaStream_t st, const float* x_t, const float* v, float alpha, float sigma, long long n,
                       float* x_pred) {
  last_launch_status = launch_k(dmd_update_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, st, x_t, v, alpha, sigma, n, x_pred);
  STTS_LAUNCH_OK();
}

cudaError_t philox_normal(cudaStream_t st, const unsigned long long* seed, unsigned long long stream_id, long long n,
                          float* out) {
  last_launch_status = launch_k(philox_normal_kernel, dim3(blocks_for((n + 3) / 4, 256)), dim3(256), 0, st, seed, stream_id, n, out);
  STTS_LAUNCH_OK();
}

cudaError_t convnext_mix(cudaStream_t st, const float* x, int B, int T, int C, const float* norm_w, const float* conv_w,
                         const float* conv_b, const float* gamma, const float* ffn_norm_w, float eps, float* y,
                         bf16* a) {
  if (C < 16 || (C & (C - 1)) != 0 || C > 2048) return cudaErrorInvalidValue;
  if (C == 128 || C == 256) {
    const int TTr = 16384 / C;
    const int smem_r = ((TTr + 6) * (C + 4) + (TTr + 6) + TTr) * 4;
    static PerDeviceOnce once_r;
    const cudaError_t er = once_r.run([] {
      cudaError_t e1 = cudaFuncSetAttribute(convnext_mix_rows_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
      cudaError_t e2 = cudaFuncSetAttribute(convnext_mix_rows_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
      return e1 != cudaSuccess ? e1 : e2;
    });
    if (er != cudaSuccess) return er;
    dim3 gr((T + TTr - 1) / TTr, B);
    if (C == 128) {
      last_launch_status = launch_k(convnext_mix_rows_kernel<128>, gr, dim3(256), smem_r, st, x, T, norm_w, conv_w, conv_b, gamma, ffn_norm_w, eps, y, a);
    } else {
      last_launch_status = launch_k(convnext_mix_rows_kernel<256>, gr, dim3(256), smem_r, st, x, T, norm_w, conv_w, conv_b, gamma, ffn_norm_w, eps, y, a);
    }
    STTS_LAUNCH_OK();
  }
  static int tt_elems = 0;  // STTS_MIX_TILE=<rows*C per CTA>: tile-size experiments
  if (tt_elems == 0) {
    const char* ev = getenv("STTS_MIX_TILE");
    tt_elems = ev ? atoi(ev) : 8192;  // measured at BASELINE configs[1]: 8192 beats 16384 by 0.09 ms per step
    if (tt_elems < 2048 || tt_elems > 24576) tt_elems = 8192;
  }
  int TT = tt_elems / C;
  if (TT < 4) TT = 4;
  // C = 2048 (75 frames per utterance): the tile + the staged parameters allow one CTA per SM, and 5-row tiles give
  // 15 x B = 120 CTAs at batch 8 -- a single wave (4-row tiles: 152 CTAs, two waves)
  if (C == 2048 && TT < 5) TT = 5;
  if (TT > 512) TT = 512;
  if (TT > T) TT = T;
  const int smem = ((TT + 6) * C + (TT + 6) + ((TT + 3) & ~3) + 10 * C) * 4;
  static PerDeviceOnce once;
  {
    const cudaError_t e = once.run([] {
      return cudaFuncSetAttribute(convnext_mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    });
    if (e != cudaSuccess) return e;
  }
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  dim3 grid((T + TT - 1) / TT, B);
  last_launch_status = launch_k(convnext_mix_kernel, dim3(grid), dim3(256), smem, st, x, T, C, TT, norm_w, conv_w, conv_b, gamma, ffn_norm_w, eps, y, a);
  STTS_LAUNCH_OK();
}

cudaError_t head_conv(cudaStream_t st, const float* x, int B, int T, int C, const float* w, const float* bias,
                      float* out) {
  if (C % 4 != 0) return cudaErrorInvalidValue;
  dim3 grid((T + 255) / 256, B);
  if (C > 40) return cudaErrorInvalidValue;  // 262 x (C+4) fp32 tile must fit the default 48 KB
  last_launch_status = launch_k(head_conv_kernel, dim3(grid), dim3(256), (7 * C + 262 * (C + 4)) * 4, st, x, T, C, w, bias, out);
  STTS_LAUNCH_OK();
}

cudaError_t pack_matrix(cudaStream_t st, const float* src, int rows, int cols, float scale, int row_mode, int row_off,
                        int col_mode, int col_off, bf16* dst, int ld_dst) {
  last_launch_status = launch_k(pack_matrix_kernel, dim3(blocks_for(static_cast<long long>(rows) * cols, 256)), dim3(256), 0, st, 
      src, rows, cols, scale, row_mode, row_off, col_mode, col_off, dst, ld_dst);
  STTS_LAUNCH_OK();
}

cudaError_t pack_conv_taps(cudaStream_t st, const float* src, int O, int cin, int taps, int kp, int opg, int group_pitch,
                           bf16* dst, int ld_dst) {
  last_launch_status = launch_k(pack_conv_taps_kernel, dim3(blocks_for(static_cast<long long>(O) * cin * taps, 256)), dim3(256), 0, st, 
      src, O, cin, taps, kp, opg, group_pitch, dst, ld_dst);
  STTS_LAUNCH_OK();
}

cudaError_t pack_conv_dense_tiles(cudaStream_t st, const float* src, int taps, bf16* dst) {
  last_launch_status = launch_k(pack_conv_dense_tiles_kernel, dim3(blocks_for(static_cast<long long>(960) * 60 * taps, 256)), dim3(256), 0, st, src, taps, dst);
  STTS_LAUNCH_OK();
}

cudaError_t audio_stem_conv(cudaStream_t st, const float* audio, int B, int N, int C, const float* w, const float* bias,
                            float* out) {
  if (C != 32) return cudaErrorInvalidValue;
  last_launch_status = launch_k(audio_stem_conv_kernel<32>, dim3(blocks_for(N, 256), B), dim3(256), 0, st, audio, N, w, bias, out);
  STTS_LAUNCH_OK();
}

cudaError_t resample(cudaStream_t st, const float* x, int B, int N, const float* bank, int K, int width, int down, int up,
                     float* out, int out_len) {
  const size_t smem = (static_cast<size_t>(31) * down + K) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  static PerDeviceOnce attr_set;
  if (smem > 48 * 1024) {
    const cudaError_t e = attr_set.run([] {
      return cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    });
    if (e != cudaSuccess) return e;
  }
  const long long frames = (static_cast<long long>(out_len) + up - 1) / up;
  last_launch_status = launch_k(resample_kernel, dim3(blocks_for(frames, 32), blocks_for(up, 8), B), dim3(256), smem, st, x, N,
                                bank, K, width, down, up, out, out_len);
  STTS_LAUNCH_OK();
}

cudaError_t pack_conv_strided(cudaStream_t st, const float* src, int O, int C, int r, bf16* dst) {
  last_launch_status = launch_k(pack_conv_strided_kernel, dim3(blocks_for(static_cast<long long>(O) * 2 * r * C, 256)), dim3(256), 0, st, src, O, C, r, dst);
  STTS_LAUNCH_OK();
}

cudaError_t pack_convtr(cudaStream_t st, const float* src, int cin, int cout, int r, bf16* dst) {
  last_launch_status = launch_k(pack_convtr_kernel, dim3(blocks_for(static_cast<long long>(cin) * cout * 2 * r, 256)), dim3(256), 0, st, src, cin, cout, r, dst);
  STTS_LAUNCH_OK();
}

cudaError_t pack_vector(cudaStream_t st, const float* src, int n, float scale, int row_mode, int off, float* dst) {
  last_launch_status = launch_k(pack_vector_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, st, src, n, scale, row_mode, off, dst);
  STTS_LAUNCH_OK();
}

cudaError_t cast_f16(cudaStream_t st, const float* src, long long n, float scale, void* dst_f16) {
  last_launch_status = launch_k(cast_f16_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, st, src, n, scale, static_cast<__half*>(dst_f16));
  STTS_LAUNCH_OK();
}

cudaError_t tile_vector(cudaStream_t st, const float* src, int n, int reps, float* dst) {
  last_launch_status = launch_k(tile_vector_kernel, dim3(blocks_for(static_cast<long long>(n) * reps, 256)), dim3(256), 0, st, src, n, reps, dst);
  STTS_LAUNCH_OK();
}

}  // namespace stts
