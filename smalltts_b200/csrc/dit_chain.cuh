// Chained DiT GEMMs: one persistent tcgen05 kernel runs up to four dependent GEMMs of a DiT block back to back
//   to_out -> w1|w3 -> w2 -> q|k|v|gate of the next block        (dit.py:197-212, :95-135, :12-25)
// with every elementwise step of the block folded into the GEMM epilogues and kernel boundaries replaced by per-row-
// block ready counters in global memory (see dit_chain.cu).  Only the attention kernel sits between two launches.
//
// What the epilogues fold (shared timestep t per launch; per-utterance t uses the generic path of engine.cu):
//   * AdaLayerNormZero (dit.py:19-25): the A operand is xb = bf16(x * (1 + scale)) written by the PRODUCING epilogue,
//     which also writes per-row partial sums (sum, sum of squares) of the fp32 x it just produced.  The consuming GEMM
//     applies LayerNorm + shift after the fact:  y = rstd * (acc - mean * cs) + b'  with the per-timestep vectors
//     cs[n] = sum_k W[n,k] (1 + scale[k]),  b'[n] = b[n] + sum_k W[n,k] shift[k]   (chain_fold_table).
//   * per-head RMSNorm + interleaved-pair RoPE + bf16 head layout (dit.py:95-108,152-173) in the q|k|v epilogue
//     (q, k, v weights are head-padded 120 -> 128 so that one 128-column tile is one head),
//   * SwiGLU (dit.py:186), tanh-gated residuals and the padded-row mask (dit.py:115-118,198,201).
#pragma once
#include "op16.cuh"
#include <cuda_runtime.h>
#include <stdint.h>

namespace stts {


// CHAIN_ATTN: the joint self | ref | text attention of a block (chain_attn.cuh) as a phase between q|k|v|gate and to_out,
// so that a whole denoiser evaluation (1 + 12 x 5 phases) is ONE launch.
enum ChainKind : int { CHAIN_QKVG = 0, CHAIN_OUT = 1, CHAIN_W13 = 2, CHAIN_W2 = 3, CHAIN_VEL = 4, CHAIN_ATTN = 5 };
constexpr int kChainMaxPhases = 64;
constexpr int kChainMaxRowBlocks = 64;  // launches with a CHAIN_ATTN phase: M <= 64 * 128 rows

constexpr int kChainD = 960, kChainH = 8, kChainHD = 120, kChainHDP = 128, kChainFF = 2400, kChainBlocks = 12;
constexpr int kChainQKVG = 3 * kChainH * kChainHDP + kChainD;  // 4032 rows: q | k | v head-padded, then the gate
constexpr int kChainW13 = 2 * kChainFF;                        // 4800 rows: w1 / w3 interleaved in groups of 16
constexpr int kChainParts = kChainD / 32;                      // 30 LayerNorm partials per row (one per 32 columns)
constexpr int kChainModLd = kChainBlocks * 6 * kChainD + 2 * kChainD;  // adaLN table of one timestep (engine.cu MOD_LD)

// Folded LayerNorm vectors of one timestep (fp32), see chain_fold_table.
constexpr int kFoldCsQ = 0;                                          // [12][4032]
constexpr int kFoldBq = kFoldCsQ + kChainBlocks * kChainQKVG;        // [12][4032]
constexpr int kFoldCs13 = kFoldBq + kChainBlocks * kChainQKVG;       // [12][4800]
constexpr int kFoldB13 = kFoldCs13 + kChainBlocks * kChainW13;       // [12][4800]
constexpr int kFoldCsV = kFoldB13 + kChainBlocks * kChainW13;        // [64]
constexpr int kFoldBv = kFoldCsV + 64;                               // [64]
constexpr int kFoldFloats = kFoldBv + 64;

struct ChainWeights {  // the 12 blocks stacked per GEMM type (device pointers, bf16 K-major like nn.Linear)
  const bf16* wqkvg = nullptr;  // [12][4032][960]
  const bf16* wo = nullptr;     // [12][960][1024]   (K head-padded 120 -> 128)
  const bf16* w13 = nullptr;    // [12][4800][960]
  const bf16* w2 = nullptr;     // [12][960][2400]
  const bf16* wvel = nullptr;   // [64][960]
  const float* bqkvg = nullptr; // [12][4032] q|k|v biases in the padded layout, gate part zero
  const float* b13 = nullptr;   // [12][4800] interleaved like w13
  const float* b2 = nullptr;    // [12][960]
  const float* bvel = nullptr;  // [64]
  const float* qn = nullptr;    // [12][8][120]
  const float* kn = nullptr;    // [12][8][120]
  const float* cos_t = nullptr; // [4096][32]
  const float* sin_t = nullptr;
};

struct ChainBuffers {  // activations of one denoiser evaluation, M = B*T rows
  float* x = nullptr;      // [M][960]  fp32 residual stream
  bf16* xb = nullptr;      // [M][960]  bf16(x * (1 + scale of the consuming LayerNorm))
  float* stats = nullptr;  // [M][30][2] partial (sum, sumsq) of x per 32-column chunk
  bf16* qkv = nullptr;     // [3][M][1024]  q | k | v, [row][head][128]; [2][3][M][1024] (by block parity, zero-initialised
                           // once) when ChainCall::qkv_db
  float* gate = nullptr;   // [M][960]  pre-sigmoid attention gate
  bf16* ob = nullptr;      // [M][1024] gated attention output (attention kernel between launches, or a CHAIN_ATTN phase)
  bf16* hb = nullptr;      // [M][2400] SwiGLU hidden
  float* vel = nullptr;    // [M][64]
  int* ready = nullptr;    // zero-initialised counters of THIS launch (chain_ready_ints of them)
  unsigned long long* trace = nullptr;  // optional role timeline [grid][64 tiles][16] (tools/trace_chain.py)
};

struct ChainAttn {  // cross-attention caches of the conditions (engine.cu stts_cond); needed by CHAIN_ATTN phases only
  int B = 0, R = 0, P = 0;                                 // utterances, max reference frames, max phonemes
  const int *ref_len = nullptr, *ph_len = nullptr;         // [B] (device)
  const bf16 *kv_ref = nullptr, *kv_text = nullptr;        // [12][2][B, R | P, 8, 128]
};

struct ChainCall {
  int M = 0, T = 0;             // rows, rows per utterance
  const int* frames = nullptr;  // [B] valid rows per utterance (device)
  const float* mod = nullptr;   // adaLN table of the timestep [kChainModLd] (gates already tanh'ed)
  const float* fold = nullptr;  // folded vectors of the timestep [kFoldFloats]
  int n_phases = 0;
  int kind[kChainMaxPhases] = {};
  int blk[kChainMaxPhases] = {};
  // q|k|v of block i go to half (i & 1) of a double buffer: with attention inside the launch, block i + 1's q|k|v GEMM of
  // one row block may run while attention items of block i still read K / V rows of that row block for other queries.
  bool qkv_db = false;
  ChainAttn attn;
};

// Number of int counters one launch needs: completed work items per (phase, row block), then (on its own 128-byte line)
// the tile-claim counter.
inline int chain_ready_ints(int M, int n_phases = 4) { return (n_phases * ((M + 127) / 128) + 31) / 32 * 32 + 64; }

// CHAIN_ATTN phases keep all keys of an utterance (self | ref | text, each padded to 16) in one 256-column accumulator.
inline bool chain_attn_fits(int T, int R, int P) { return ((T + 15) & ~15) + ((R + 15) & ~15) + ((P + 15) & ~15) <= 256; }

cudaError_t launch_dit_chain(cudaStream_t st, const ChainWeights& w, const ChainBuffers& b, const ChainCall& c);

// fold[...] of one timestep from its adaLN table (run once per timestep, cached by the engine).
cudaError_t chain_fold_table(cudaStream_t st, const ChainWeights& w, const float* mod, float* fold);

// Entry of the chain: xb = bf16(x * (1 + scale)) and the LayerNorm partials of x (the first block's input comes from
// the input embedding, whose GEMM epilogue does not write them).
cudaError_t chain_stats_cast(cudaStream_t st, const float* x, int M, const float* scale, bf16* xb, float* stats);

// Head-padded row packing for the stacked q|k|v|gate weight: dst[(r / 120) * 128 + r % 120 + row_off, c] = bf16(src[r, c])
cudaError_t pack_rows_headpad(cudaStream_t st, const float* src, int rows, int cols, int row_off, bf16* dst, int ld_dst);
cudaError_t pack_vec_headpad(cudaStream_t st, const float* src, int n, int off, float* dst);

}  // namespace stts
