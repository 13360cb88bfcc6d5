// Non-GEMM kernels of the engine (declarations).  All launch on the given stream and return the
// CUDA launch status.  Layouts are channels-last: activations [B, T, C] row-major.
#pragma once
#include "op16.cuh"
#include <cuda_runtime.h>
#include <stdint.h>

namespace stts {


// ---- normalisation producing the bf16 A operand of the next GEMM
// LayerNorm(no affine, eps) * (1 + scale[b]) + shift[b]   (dit.py:24,199 / :38).  scale/shift: [B or 1, ld_mod].
cudaError_t ln_mod_bf16(cudaStream_t st, const float* x, int rows, int rows_per_batch, int dim, const float* scale,
                        const float* shift, int ld_mod, float eps, bf16* out);
// x * rsqrt(mean(x^2) + eps) * w   (dit.py:52 1-D weight; style.py:100-105 block norms)
cudaError_t rms_norm_bf16(cudaStream_t st, const float* x, int rows, int dim, const float* w, float eps, bf16* out);

// ---- head split: fp32 [rows, ld_in] (columns src_off + h*hd + d) -> bf16 [rows, heads, hd_pad], with optional
// per-head RMSNorm weight [heads, hd] (dit.py:52-53) and optional interleaved-pair RoPE on the first `rot` dims
// using position (row % rows_per_batch) (dit.py:152-173 / style.py:21-25).  cos/sin tables: [max_pos, rot/2].
cudaError_t head_split_bf16(cudaStream_t st, const float* in, int ld_in, int src_off, int rows, int rows_per_batch,
                            int heads, int hd, int hd_pad, const float* norm_w, float eps, int rot, const float* cos_t,
                            const float* sin_t, bf16* out);

// q, k, v of one fused projection in ONE launch: sources at column offsets 0, stride_off, 2*stride_off; q and k get
// per-head RMSNorm + RoPE, v is copied (head-padded).
cudaError_t head_split_qkv_bf16(cudaStream_t st, const float* in, int ld_in, int stride_off, int rows, int rows_per_batch,
                                int heads, int hd, int hd_pad, const float* q_norm, const float* k_norm, float eps,
                                int rot, const float* cos_t, const float* sin_t, bf16* q, bf16* k, bf16* v);

// Cross K/V caches of all `layers` DiT blocks from one fused projection (dit.py:80-93): in fp32 [rows, ld_in] with
// layer l's k at columns 2*l*heads*hd and v right after; k gets per-head RMSNorm with knorm_all[l] ([layers, heads, hd]),
// v is copied; output y = 2*l + {0,1} goes to cache + y*out_stride as bf16 [rows, heads, hd_pad].
cudaError_t kv_split_bf16(cudaStream_t st, const float* in, int ld_in, int rows, int layers, int heads, int hd,
                          int hd_pad, float eps, const float* knorm_all, bf16* cache, long long out_stride);

// ---- attention over up to three key segments (self | ref | text), prefix-valid lengths per batch.
struct AttnSeg {
  const bf16* k = nullptr;   // [B, n_max, H, hd_pad]
  const bf16* v = nullptr;
  const int* len = nullptr;  // [B] valid prefix length (nullptr -> n_max)
  int n_max = 0;
};
// q: [B, tq, H, hd_pad]; out: bf16 [B*tq, H*hd_pad]; gate (optional): fp32 pre-sigmoid [B*tq, ld_gate] at column
// gate_off + h*hd + d (dit.py:111-115); softmax scale = 1/sqrt(hd).
cudaError_t attention_bf16(cudaStream_t st, const bf16* q, int B, int tq, int H, int hd, int hd_pad, const AttnSeg* segs,
                           int nseg, const float* gate, int ld_gate, int gate_off, bf16* out);

// ---- small dense layers on <= 16 rows (time embedding, adaLN tables): fp32 weights, warp per output.
// y[r, n] = post( sum_k pre(x[r,k]) * W[n,k] + b[n] ); pre/post: 0 none, 1 SiLU; tanh applied to output chunk
// (n / chunk) when bit set in tanh_chunks.
cudaError_t gemv_rows(cudaStream_t st, const float* x, int rows, int k, const float* w, const float* b, int n, int pre,
                      int post, int chunk, unsigned tanh_chunks, float* y, int ld_y);
// sinusoidal timestep features (model.py:23-29): out [rows, 256]
cudaError_t time_features(cudaStream_t st, const float* t, int rows, float* out);

// ---- text embedding gather: ids int64 [rows] -> fp32 [rows, dim]
cudaError_t embed_gather(cudaStream_t st, const long long* ids, int rows, const float* table, int vocab, int dim,
                         float* out);
cudaError_t cast_bf16(cudaStream_t st, const float* in, long long n, bf16* out);

// ---- sampler math (infer/onnx.py:105,125): all [n] fp32
// x_t = alpha*x_pred + sigma*noise ; writes fp32 and bf16 copies
cudaError_t noise_mix(cudaStream_t st, const float* x_pred, const float* noise, float alpha, float sigma, long long n,
                      float* x_t, bf16* x_t_bf16);
// x_pred = alpha*x_t - sigma*v
cudaError_t dmd_update(cudaStream_t st, const float* x_t, const float* v, float alpha, float sigma, long long n,
                       float* x_pred);
// ---- teacher sampler (config 5): x_t shared by the three CFG branches, CFG combine + DDIM step
cudaError_t repeat_cast_bf16(cudaStream_t st, const float* in, long long n, int reps, bf16* out);
// v = v_c + s_text (v_c - v_no_text) + s_spk (v_c - v_no_spk) with v3 = [v_c | v_no_text | v_no_spk] (distill.py:97-103);
// x <- ca * x + cb * v
cudaError_t cfg_ddim_update(cudaStream_t st, float* x, const float* v3, long long n, float s_text, float s_spk, float ca,
                            float cb);
// standard normal noise, Philox4x32-10 + Box-Muller, counter = element index; seed read from device memory
cudaError_t philox_normal(cudaStream_t st, const unsigned long long* seed, unsigned long long stream_id, long long n,
                          float* out);

// ---- vocoder ConvNeXt token mixer (hf:284-292) fused with the FFN pre-norm (hf:295):
//   y = x + gamma * (dwconv7_causal(rmsnorm(x; norm_w)) + conv_b) ;  a = bf16(rmsnorm(y; ffn_norm_w))
// x, y: fp32 [B, T, C]; conv_w: [C, 7]
cudaError_t convnext_mix(cudaStream_t st, const float* x, int B, int T, int C, const float* norm_w, const float* conv_w,
                         const float* conv_b, const float* gamma, const float* ffn_norm_w, float eps, float* y,
                         bf16* a);
// Whole ConvNeXt layer (token mixer + FFN, hf:284-297) in one persistent tcgen05 kernel for C = 32 / 64: reads x once,
// writes out once (fp32), optional bf16 copy of out.  w1: bf16 [4C, C], w2: FP16 [C, 4C] (the hidden activation is
// produced in fp16).  x and out must not alias.
cudaError_t convnext_fused(cudaStream_t st, const float* x, int B, int T, int C, const float* norm_w,
                           const float* conv_w, const float* conv_b, const float* gamma, const float* ffn_norm_w,
                           const bf16* w1, const float* b1, const void* w2_f16, const float* b2, const float* ffn_gamma,
                           float eps, float* out, bf16* out_bf16);
// ConvNeXt feed-forward (hf:293-297) for C = 128 in one kernel; the 4C-wide hidden activation never leaves the SM:
//   out = y + ffn_gamma * (W2 gelu(W1 a + b1) + b2),  a: bf16 [M, C] (FFN pre-norm output), y: fp32 [M, C] residual,
// w1: bf16 [4C, C], w2: FP16 [C, 4C] pre-scaled by 0.5.  out must not alias y.
cudaError_t ffn_fused(cudaStream_t st, const bf16* a, const float* y, long long M, int C, const bf16* w1, const float* b1,
                      const void* w2_f16, const float* b2, const float* ffn_gamma, float* out, bf16* out_bf16);
// vocoder head: causal Conv1d(C -> 1, k=7) (hf:484-489).  x fp32 [B, T, C], w [C, 7] -> out [B, T]
cudaError_t head_conv(cudaStream_t st, const float* x, int B, int T, int C, const float* w, const float* bias,
                      float* out);

// codec encoder stem: causal Conv1d(1 -> C, k=7) (hf:300-312).  audio fp32 [B, N], w [C, 1, 7] -> out fp32 [B, N, C]
cudaError_t audio_stem_conv(cudaStream_t st, const float* audio, int B, int N, int C, const float* w, const float* bias,
                            float* out);

// polyphase FIR resampler (infer/utils.py:7-23 == torchaudio sinc_interp_kaiser): x fp32 [B, N] -> out fp32 [B, out_len],
// out[n*up + j] = sum_k bank[j, k] * x[n*down + k - width];  bank fp32 [up, K], K = 2*width + down
cudaError_t resample(cudaStream_t st, const float* x, int B, int N, const float* bank, int K, int width, int down, int up,
                     float* out, int out_len);

// ---- weight packing (run once at stts_finalize_weights)
enum PackRow : int { ROW_PLAIN = 0, ROW_INTERLEAVE16_LO = 1, ROW_INTERLEAVE16_HI = 2, ROW_GROUPPAD_60_64 = 3 };
enum PackCol : int { COL_PLAIN = 0, COL_HEADPAD_120_128 = 1 };
// dst[rowmap(r) + row_off, colmap(c) + col_off] = bf16(scale * src[r, c])
cudaError_t pack_matrix(cudaStream_t st, const float* src, int rows, int cols, float scale, int row_mode, int row_off,
                        int col_mode, int col_off, bf16* dst, int ld_dst);
// Conv1d weight [O, Cin, taps] -> dst[(o / opg) * group_pitch + o % opg, tap * kp + c]
cudaError_t pack_conv_taps(cudaStream_t st, const float* src, int O, int cin, int taps, int kp, int opg, int group_pitch,
                           bf16* dst, int ld_dst);
// Grouped Conv1d(960,960,k,groups=16) weight [960, 60, taps] -> 15 dense 64-column output tiles, each reading the two
// adjacent 64-padded input groups it can touch: dst[o, tap*128 + (o/60 - o/64)*64 + c] = w[o, c, tap]  (dst [960, taps*128])
cudaError_t pack_conv_dense_tiles(cudaStream_t st, const float* src, int taps, bf16* dst);
// strided causal Conv1d weight [O, C, 2r] (hf:181-216, k = 2r, stride r) -> dst[o, tap*r*C + j*C + c] = w[o, c, tap*r + j]
// (two taps over the input regrouped r rows at a time: [T, C] viewed as [T/r, r*C])
cudaError_t pack_conv_strided(cudaStream_t st, const float* src, int O, int C, int r, bf16* dst);
// ConvTranspose1d weight [Cin, Cout, 2r] -> dst[j*Cout + o, tap*Cin + c] = w[c, o, j + tap*r]
cudaError_t pack_convtr(cudaStream_t st, const float* src, int cin, int cout, int r, bf16* dst);
// fp32 vector helpers: dst[map(i) + off] = scale * src[i]
cudaError_t pack_vector(cudaStream_t st, const float* src, int n, float scale, int row_mode, int off, float* dst);
cudaError_t tile_vector(cudaStream_t st, const float* src, int n, int reps, float* dst);  // dst[j*n+i] = src[i]
cudaError_t cast_f16(cudaStream_t st, const float* src, long long n, float scale, void* dst_f16);  // dst = fp16(scale*src)


}  // namespace stts
