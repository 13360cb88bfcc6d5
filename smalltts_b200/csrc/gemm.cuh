// Multi-tap tcgen05 GEMM: the one tensor-core workhorse of the engine.
//
//   C[b, t, g*out_group_cols + n] = epilogue( sum_{tap} sum_{k<K} A[b, t + shift(tap), g*a_group_koff + k]
//                                                          * W[g*w_group_rows + n, tap*Kp + k] )
//
// A is a channels-last bf16 activation [B, T, ldA] read through a 3-D TMA tensor map whose out-of-bounds
// rows/columns read as zero.  That single property turns every convolution on the hot path into this GEMM:
//   * nn.Linear                       : taps=1, B=1, T=rows (flattened batch)
//   * causal Conv1d k=7 (vocoder stem): taps=7, shift = tap-6        (hf:181-216)
//   * causal ConvTranspose1d k=2r,s=r : taps=2, shift = 0,-1, N=r*Cout (hf:219-260; trim is implicit)
//   * grouped Conv1d k=31,g=16,pad=15 : taps=31, shift = tap-15, groups=16 (dit.py:223-236), channels of each
//                                       group padded 60 -> 64 so that group offsets are 16-byte aligned
// W is bf16 [rows, taps*Kp] K-major (nn.Linear's own [out,in] layout), Kp = K rounded up to 64.
#pragma once
#include <cuda.h>
#include "op16.cuh"
#include <cuda_runtime.h>
#include <stdint.h>

namespace stts {

enum GemmAct : int {
  ACT_NONE = 0,
  ACT_GELU = 1,      // erf GELU (hf ACT2FN["gelu"])
  ACT_MISH = 2,      // x*tanh(softplus(x)) (dit.py:226-229)
  ACT_SIGMOID = 3,
  ACT_SWIGLU16 = 4,  // columns interleaved in 16s: [w1 x16 | w3 x16]; out = silu(a)*b, N/2 output columns
  ACT_SILU = 5,
};

struct GemmShape {
  int B = 1, T = 0;  // B batches of T rows
  int N = 0;         // valid output columns per group
  int K = 0;         // reduction length per tap
  int taps = 1, tap_shift0 = 0, tap_step = 1;
  // groups: A columns start at g*a_group_koff (must be a multiple of 8 elements: TMA needs 16-byte aligned box
  // starts), W rows at g*w_group_rows, output columns at g*out_group_cols, residual columns at g*res_group_cols
  int groups = 1, a_group_koff = 0, w_group_rows = 0, out_group_cols = 0, res_group_cols = 0;
  int ab_f16 = 0;  // A and W hold fp16 (not bf16) values: selects the fp16 input format of tcgen05.mma kind::f16
  // Split-K: every output tile is computed by `splits` work items, each over a contiguous part of the taps*K reduction
  // (wave quantisation: 160 tiles on 148 SMs, or 80 tiles of 128 k-iterations, leave half of the chip idle).  Every part
  // parks its fp32 accumulator in GemmEpi::split_scratch; the part that arrives LAST at the tile's counter adds the
  // parts in part order (so the result does not depend on who was last) and runs the epilogue.  Nobody waits for
  // anybody: no residency assumption.  Single-CTA kernel only.
  int splits = 1;
};

struct GemmEpi {
  const float* bias = nullptr;      // [cols]
  int act = ACT_NONE;
  int rows_per_batch = 0;           // >0: batch index for row_len/rowgate = row / rows_per_batch (flattened A)
  const int* row_len = nullptr;     // [B]; rows t >= row_len[b] are written as 0 (after activation)
  int mask_bf16_only = 0;           // apply row_len masking to the bf16 output only
  int gelu2_f16 = 0;                // ACT_GELU only: write 2*gelu(x) as FP16 into out_bf16 (packed half2 math); the
                                    // consumer GEMM runs with ab_f16 = 1 and weights pre-scaled by 0.5
  const float* colscale = nullptr;  // [cols]
  const float* rowgate = nullptr;   // [B, ld_gate]
  int ld_gate = 0;
  const float* residual = nullptr;  // [B*T, ld_res] fp32 (may alias out_f32)
  int ld_res = 0;
  float* out_f32 = nullptr;
  bf16* out_bf16 = nullptr;
  int ld_out = 0;
  // split-K only (GemmShape::splits > 1): fp32 scratch of gemm_split_scratch_floats() elements and zero-initialised
  // counters (8 per output tile; the kernel leaves them zero)
  float* split_scratch = nullptr;
  int* split_counters = nullptr;
};

// Scratch elements / counters a split-K launch of this shape needs (block_n as passed to launch_gemm).
long long gemm_split_scratch_floats(const GemmShape& s, int block_n);
long long gemm_split_counters(const GemmShape& s, int block_n);

struct GemmA {
  const bf16* ptr;
  int cols;  // valid columns (tensor-map extent; reads beyond are zero)
  int ld;    // row pitch in elements (multiple of 8)
};
struct GemmW {
  const bf16* ptr;
  int rows;  // tensor-map extent (rows beyond read as zero)
  int ld;    // row pitch in elements = taps*Kp (multiple of 8)
};

// `block_n | kGemmPairFlag` (block_n 128 or 256) runs the CTA-pair variant: clusters of two CTAs compute 256 x block_n
// tiles with cta_group::2 MMAs (ACT_NONE and the fp16 2*gelu epilogue only).
constexpr int kGemmPairFlag = 0x1000;

// Returns cudaSuccess or the first CUDA error; `block_n` in {32, 64, 128, 256}.
cudaError_t launch_gemm(cudaStream_t stream, int block_n, const GemmA& a, const GemmW& w, const GemmShape& s,
                        const GemmEpi& e);


}  // namespace stts
