// Joint self | ref | text attention (dit.py:110-119,131-135) as a work item of the chained DiT kernel (dit_chain.cu), on
// tcgen05: one item = 128 query rows of one (utterance, head) against ALL keys of the utterance (at most 256 after
// padding every segment to a multiple of 16 -- the one-launch schedule is only chosen for such shapes):
//   S[128 x Nk]  = Q K^T          one UMMA chain, M 128 x N Nk x K 128, accumulator in TMEM columns [0, 256)
//   P            = exp2((S - rowmax) * log2e / sqrt(120)), masked, bf16, written by the epilogue warps to shared memory
//                  in the K-major 128-byte-swizzle operand layout
//   O[128 x 128] = P V            second UMMA chain, K = Nk; V stays in its natural [key][dim] layout and is consumed as an
//                  MN-major B operand; accumulator in TMEM columns [256, 384)
//   out          = O / rowsum * sigmoid(gate)
// The item borrows the (drained) operand ring of the GEMM phases:  Q 32 KB | K 64 KB (later P) | V 64 KB | gates 32 KB.
// K / V boxes are 16 keys x 64 dims (2 KB) taken from the q|k|v buffer of the launch or from the cross-attention caches of
// the conditions, so the virtual key order is [self | ref | text], each padded to 16 keys; padded keys are masked.
#pragma once
#include <cuda.h>
#include <math.h>

#include "dit_chain.cuh"
#include "ptx.cuh"

namespace stts {
namespace chain_attn {

constexpr int kRows = 128;                  // query rows per item (UMMA M)
constexpr int kMaxKeys = 256;               // virtual keys per utterance (UMMA N of the score MMA, TMEM columns)
constexpr int kHD = 128;                    // padded head dim
constexpr int kQBytes = 2 * kRows * 128;    // [2 dim halves][128 rows][128 B]
constexpr int kKBytes = 2 * kMaxKeys * 128; // [2 dim halves][256 keys][128 B]; P reuses it: [4 key blocks][128 rows][128 B]
constexpr int kVBytes = kKBytes;
constexpr int kQOff = 0, kKOff = kQBytes, kPOff = kKOff, kVOff = kKOff + kKBytes;
constexpr int kGateOff = kVOff + kVBytes;    // sigmoid(gate) of the item's rows, fp16 [128 rows][128 dims]
constexpr int kGateBytes = kRows * kHD * 2;
constexpr int kSmemBytes = kQBytes + kKBytes + kVBytes + kGateBytes;
constexpr int kSCol = 0, kOCol = 256;       // TMEM columns of the two accumulators
constexpr int kBarriers = 5;                // qk_full (Q by the A producer + K by the W producer), v_full, s_full, p_full, o_full
static_assert(kRows * kMaxKeys * 2 <= kKBytes, "P must fit in the K buffer");

struct Maps {
  CUtensorMap q;           // q|k|v buffer(s) [(2 x) 3*M rows][8 heads][128]: box 64 dims x 1 head x 128 rows
  CUtensorMap kv_self[2];  // one half of the buffer each [3*M rows], box 64 x 1 x 16 rows (k rows at M + ., v at 2M + .)
  CUtensorMap ref;         // cross caches [12][2][B*R rows][8][128], box 64 x 1 x 16
  CUtensorMap text;        // [12][2][B*P rows][8][128]
};

// virtual key layout of one utterance: segment s occupies [e(s-1), e(s-1) + pad16(len_s))
struct Keys {
  int len0, len1, len2, e0, e1, e2, ng;  // ng = 16-key groups (<= 16)
};
__device__ __forceinline__ Keys keys_of(const ChainCall& c, int b) {
  Keys k;
  k.len0 = min(__ldg(c.frames + b), c.T);
  k.len1 = min(__ldg(c.attn.ref_len + b), c.attn.R);
  k.len2 = min(__ldg(c.attn.ph_len + b), c.attn.P);
  k.e0 = (k.len0 + 15) & ~15;
  k.e1 = k.e0 + ((k.len1 + 15) & ~15);
  k.e2 = k.e1 + ((k.len2 + 15) & ~15);
  k.ng = k.e2 >> 4;
  return k;
}
__device__ __forceinline__ bool key_valid(const Keys& k, int kv) {
  return kv < k.e0 ? kv < k.len0 : (kv < k.e1 ? kv - k.e0 < k.len1 : kv - k.e1 < k.len2);
}

// One lane: the K box and the V box of 16-key group g, dim half hf.  self: taken from the q|k|v half `par` of this launch.
__device__ __forceinline__ void load_group(const Maps& maps, const ChainCall& c, const Keys& k, int blk, int par, int b, int h,
                                           int g, int hf, uint8_t* ring, uint64_t* k_bar, uint64_t* v_bar) {
  const int vk = g * 16;
  const CUtensorMap* mp;
  long long rk, rv;
  if (vk >= k.e1) {
    mp = &maps.text;
    rk = static_cast<long long>(2 * blk) * c.attn.B * c.attn.P + static_cast<long long>(b) * c.attn.P + (vk - k.e1);
    rv = rk + static_cast<long long>(c.attn.B) * c.attn.P;
  } else if (vk >= k.e0) {
    mp = &maps.ref;
    rk = static_cast<long long>(2 * blk) * c.attn.B * c.attn.R + static_cast<long long>(b) * c.attn.R + (vk - k.e0);
    rv = rk + static_cast<long long>(c.attn.B) * c.attn.R;
  } else {
    mp = &maps.kv_self[par];
    rk = static_cast<long long>(c.M) + static_cast<long long>(b) * c.T + vk;
    rv = rk + c.M;
  }
  ptx::tma_load_3d(ring + kKOff + hf * (kMaxKeys * 128) + g * 2048, mp, k_bar, hf * 64, h, static_cast<int>(rk));
  ptx::tma_load_3d(ring + kVOff + hf * (kMaxKeys * 128) + g * 2048, mp, v_bar, hf * 64, h, static_cast<int>(rv));
}

// Shared-memory descriptor of an MN-major B operand with 128-byte swizzle (cute/atom/mma_traits_sm100.hpp, canonical layout
// ((8,n),(8,k)) : ((1,LBO),(8,SBO)) in 16-byte units): rows = K index (keys) at a 128-byte pitch, 8-row groups SBO = 1024
// bytes apart; a row holds 64 elements of the MN index (dims), the next 64 are LBO bytes away (the other dim half).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
constexpr uint32_t kIdescBMajorMN = 1u << 16;  // instruction descriptor: B is MN-major (cute/arch/mma_sm100_desc.hpp)

}  // namespace chain_attn
}  // namespace stts
