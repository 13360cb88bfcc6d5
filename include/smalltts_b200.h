/*
 * smalltts_b200 -- C ABI of the B200-native engine for the smalltts synthesize hot path.
 *
 * This is the drop-in boundary.  The reference reaches this path through three opaque
 * onnxruntime sessions plus host-side numpy sampling:
 *
 *   reference call site                                        replaced by
 *   ---------------------------------------------------------  -----------------------------------------
 *   ort.InferenceSession(path, ...)  infer/onnx.py:21-24,60-63  stts_create + stts_load_weight* + stts_finalize_weights
 *     (Rust: Session::builder()...   server/src/pipeline.rs:16-20,40-48)
 *   cond_enc.run(None, feed)         infer/onnx.py:91-96        stts_encode_conditions
 *     (Rust: Pipeline::cond_encode   pipeline.rs:122-143)
 *   denoiser.run(None, feed)[0]      infer/onnx.py:107-124      stts_denoise_step
 *     (Rust: Pipeline::denoise       pipeline.rs:145-166)
 *   the 4-step numpy loop            infer/onnx.py:98-125       stts_sample           (loop stays on device)
 *     (Rust:                         pipeline.rs:81-93)
 *   codec_dec.run(None, feed)[0]     infer/onnx.py:127-128      stts_decode
 *     (Decoder.decode codec/onnx.py:42-53; Rust pipeline.rs:168-174)
 *   SmallTTS.synthesize              infer/onnx.py:68-129       stts_synthesize       (all of the above, one call)
 *     (Rust: Pipeline::synthesize_timed pipeline.rs:60-112)
 *   Timing{...}                      pipeline.rs:29-37          stts_get_timings
 *
 * Conventions
 *   - extern "C", plain pointers and sizes.  Every function returns 0 on success or a negative stts_status;
 *     stts_last_error() gives the message (per engine; pass NULL for creation failures).
 *   - The engine owns its CUDA device memory (weights, workspaces, conditions) and one CUDA stream.  A handle
 *     is NOT thread-safe: one call in flight per engine (the reference serialises with a mutex, main.rs:25).
 *   - Data pointers may be host or device memory; the `mem` argument says which (STTS_MEM_HOST / STTS_MEM_DEVICE).
 *     Host buffers are copied with cudaMemcpyAsync on the engine stream (pinned buffers make that truly async);
 *     the call returns after the stream has been synchronised, so outputs are ready on return.
 *   - Control metadata -- every length array (ref_len, ph_len, frames) and every timestep array -- is ALWAYS host
 *     memory, whatever `mem` says; only tensors (latents, ids, noise, audio) follow `mem`.
 *   - Ragged batches are padded: utterance b owns rows [0, len[b]) of its [*, max, C] slab.  Masks of the
 *     reference operators (bool tensors) are prefix masks built from these lengths (infer/onnx.py:88-99).
 *   - There is no CPU fallback: stts_create fails if the device is not compute capability 10.x.
 */
#ifndef SMALLTTS_B200_H_
#define SMALLTTS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STTS_SAMPLE_RATE 24000 /* infer/onnx.py:11 */
#define STTS_HOP_SIZE 3200     /* infer/onnx.py:12 */
#define STTS_LATENT_DIM 64
#define STTS_NUM_STEPS 4 /* infer/onnx.py:13 */

typedef enum stts_status {
  STTS_OK = 0,
  STTS_ERR_INVALID = -1,    /* bad argument / shape */
  STTS_ERR_CUDA = -2,       /* CUDA runtime or kernel failure */
  STTS_ERR_NO_DEVICE = -3,  /* no sm_100 device: there is no fallback */
  STTS_ERR_WEIGHTS = -4,    /* missing / mis-shaped weight, or not finalized */
  STTS_ERR_OOM = -5
} stts_status;

typedef enum stts_mem { STTS_MEM_HOST = 0, STTS_MEM_DEVICE = 1 } stts_mem;

typedef struct stts_engine stts_engine;
typedef struct stts_cond stts_cond;

typedef struct stts_config {
  int device;        /* CUDA ordinal */
  int reserved[7];   /* zero */
} stts_config;

/* Stage timings in milliseconds of the last stts_synthesize call, CUDA-event timed; same split as the
 * reference server's Timing (pipeline.rs:29-37). codec_enc_ms is always 0 (encoder not on this path). */
typedef struct stts_timing {
  float codec_enc_ms, cond_enc_ms, denoise_ms, codec_dec_ms, total_ms;
} stts_timing;

int stts_create(const stts_config* cfg, stts_engine** out);
void stts_destroy(stts_engine* e);
/* A second handle on the same device that SHARES the finalized weights of `src` (no copy) and owns its own streams,
 * per-shape plans / CUDA graphs and timers: two handles may run calls concurrently from two host threads, which is how
 * a server keeps two batches in flight (the reference serialises on one mutex, main.rs:25).  Destroy every clone before
 * `src`; loading or finalizing weights through a clone is not allowed. */
int stts_engine_clone(stts_engine* src, stts_engine** out);
const char* stts_last_error(const stts_engine* e);

/* Weights: fp32 tensors under the reference's own state-dict names.
 * model: 0 = DiTModel (models/backbone/model.py:33-54), 1 = VibeVoice acoustic-tokenizer decoder,
 *        2 = VibeVoice acoustic-tokenizer encoder (optional; only stts_encode_audio needs it).
 * `data` is host fp32, row-major, `ndim` <= 4.  stts_finalize_weights packs them for the kernels (bf16,
 * K-major, tap-major conv layouts) and validates that every tensor of both models is present. */
int stts_load_weight(stts_engine* e, int model, const char* name, const float* data, int ndim, const int64_t* shape);
int stts_finalize_weights(stts_engine* e);

/* == condition_encoder.onnx.  ref [B,R,64] f32, ref_len [B] i64, phonemes [B,P] i64, ph_len [B] i64
 * (phonemes_mask[b, p] = p < ph_len[b]).  The result lives on the device and may be reused for any number of
 * stts_denoise_step / stts_sample calls with the same batch (voice cloning: one condition, many prompts). */
int stts_encode_conditions(stts_engine* e, const float* ref, const int64_t* ref_len, const int64_t* phonemes,
                           const int64_t* ph_len, int B, int R, int P, int mem, stts_cond** out);
void stts_cond_free(stts_engine* e, stts_cond* c);
/* Test hook: copy block `layer`'s cached cross K/V out as fp32 [B,8,N,120] like the reference's
 * project_cross_kv (dit.py:88-93); which: 0 k_ref, 1 v_ref, 2 k_text, 3 v_text. Host destination. */
int stts_cond_read_kv(stts_engine* e, const stts_cond* c, int layer, int which, float* dst);

/* == denoiser.onnx: velocity [B,T,64] for x_t [B,T,64], frames [B] i64 (mask[b,t] = t < frames[b]), t [B] f32 (host). */
int stts_denoise_step(stts_engine* e, const stts_cond* c, const float* x_t, const int64_t* frames, const float* t,
                      int B, int T, int mem, float* velocity);

/* DMD re-noising loop, on device.  timesteps: host [steps] (NULL -> linspace(1,0,steps), infer/onnx.py:102).
 * noise: [steps,B,T,64] (NULL -> on-device Philox normal stream keyed by `seed`).  out_latents: [B,T,64]. */
int stts_sample(stts_engine* e, const stts_cond* c, const int64_t* frames, int B, int T, int steps,
                const float* timesteps, const float* noise, uint64_t seed, int mem, float* out_latents);

/* Teacher sampler for the latency/quality sweep (BASELINE config 5).  The reference ships no teacher inference
 * script; this is the sampler its distillation code implies: 3-way classifier-free guidance as built by get_x_pred
 * (scripts/train/dmd2/distill.py:60-134: batch rows [cond | text dropped | speaker dropped], velocity =
 * v_c + s_text (v_c - v_no_text) + s_spk (v_c - v_no_spk)) and a deterministic DDIM walk over
 * t = linspace(1, 0, steps + 1) with the v-prediction identities of train/utils.py:54-67.
 * cond3: stts_encode_conditions of the 3*B-row batch [ref,len,ids,plen | ref,len,0,0 | ref,0,ids,plen] (a dropped
 * condition is a zero length).  noise: [B,T,64] (the start x_1) or NULL for on-device Philox(seed).
 * out_latents: [B,T,64]. */
int stts_sample_teacher(stts_engine* e, const stts_cond* cond3, const int64_t* frames, int B, int T, int steps,
                        float cfg_scale_text, float cfg_scale_speaker, const float* noise, uint64_t seed, int mem,
                        float* out_latents);

/* == codec/decoder.onnx: latents [B,T,64] -> audio [B, T*3200] (the reference's (B,1,T*3200)). */
int stts_decode(stts_engine* e, const float* latents, int B, int T, int mem, float* audio);

/* Codec encoder: replaces `Encoder.encode` (codec/onnx.py:56-75 == assets/codec/encoder.onnx, the clone path of
 * scripts/infer/clone.py:36; Rust twin pipeline.rs codec_enc stage).  audio: fp32 [B, N] mono 24 kHz, N a positive
 * multiple of 3200 (the encoder is causal and floors, so trimming a tail shorter than one hop does not change any
 * latent); latents: fp32 [B, N/3200, 64] (the VAE mean).  Needs the encoder tensors (model index 2 of
 * stts_load_weight, HF VibeVoiceAcousticTokenizerEncoderModel.state_dict() names) loaded before stts_finalize_weights.
 * Sets stts_timing.codec_enc_ms. */
int stts_encode_audio(stts_engine* e, const float* audio, int B, int N, int mem, float* latents);

/* High-quality resampler of the clone path: replaces `resample_hq` (infer/utils.py:7-23 == torchaudio
 * Resample(resampling_method="sinc_interp_kaiser", lowpass_filter_width=1024, rolloff=0.94,
 * beta=14.769656459379492), called by scripts/infer/clone.py:32 before Encoder.encode).  audio: fp32 [B, N] at
 * sr_from; out: fp32 [B, stts_resample_length(N, sr_from, sr_to)] at sr_to.  The polyphase filter bank of a rate pair
 * is built on first use and cached in the engine.  sr_from == sr_to copies.  Needs no model weights. */
int64_t stts_resample_length(int N, int sr_from, int sr_to); /* ceil(N * sr_to / sr_from); -1 on bad arguments */
int stts_resample(stts_engine* e, const float* audio, int B, int N, int sr_from, int sr_to, int mem, float* out);

/* == SmallTTS.synthesize for a padded batch; intermediates never leave the device. audio: [B, T*3200]. */
int stts_synthesize(stts_engine* e, const float* ref, const int64_t* ref_len, const int64_t* phonemes,
                    const int64_t* ph_len, const int64_t* frames, int B, int R, int P, int T, int steps,
                    const float* timesteps, const float* noise, uint64_t seed, int mem, float* audio);

int stts_get_timings(const stts_engine* e, stts_timing* out);
/* Kernels launched by this library in this process so far (bench.py's gpu_launches). */
uint64_t stts_launch_count(void);
/* Per-call device timing of the dominant vocoder kernels (ms, summed over the last stts_decode/stts_synthesize):
 * which: 0 = HBM-bound tail (stages with C <= 128 + head), 1 = tensor-bound front (stem .. C >= 256). */
float stts_last_vocoder_ms(const stts_engine* e, int which);

/* Device-side stopwatch on the engine's stream (CUDA events): start records, stop records + synchronises and
 * returns the elapsed milliseconds of everything the engine ran in between. */
int stts_timer_start(stts_engine* e);
int stts_timer_stop(stts_engine* e, float* ms);

/* Pinned host memory helpers for callers that want truly asynchronous copies. */
void* stts_host_alloc(size_t bytes);
void stts_host_free(void* p);

/* ---- kernel-level test hooks (used by tests/ only; all pointers are DEVICE memory) ---- */
/* on != 0: the stts_test_* hooks below return without synchronising the engine stream (micro-benchmarks bracket many
 * launches with stts_timer_start/stop). */
int stts_test_set_async(stts_engine* e, int on);

int stts_test_gemm(stts_engine* e, int block_n, const void* a_bf16, int B, int T, int a_cols, int a_ld,
                   const void* w_bf16, int w_rows, int w_ld, int N, int K, int taps, int tap_shift0, int tap_step,
                   int groups, int a_group_koff, int w_group_rows, int out_group_cols, const float* bias, int act,
                   const int32_t* row_len, int rows_per_batch, int mask_bf16_only, const float* colscale, const float* rowgate, int ld_gate,
                   const float* residual, int ld_res, float* out_f32, void* out_bf16, int ld_out);
/* Split-K linear (gemm.cuh GemmShape::splits): out = epi(A[M,K] W[N,K]^T) with `splits` parts per tile; runs the launch
 * twice on one counter buffer (the kernel must leave its counters zero). */
int stts_test_gemm_split(stts_engine* e, int block_n, int splits, const void* a_bf16, int M, int K, const void* w_bf16, int N,
                         const float* bias, int gelu2_f16, const float* colscale, const float* residual, float* out_f32,
                         void* out_bf16);
int stts_test_attention(stts_engine* e, const void* q, int B, int tq, int H, int hd, int hd_pad, const void* k0,
                        const void* v0, const int32_t* len0, int n0, const void* k1, const void* v1,
                        const int32_t* len1, int n1, const float* gate, int ld_gate, void* out);
int stts_test_convnext_mix(stts_engine* e, const float* x, int B, int T, int C, const float* norm_w,
                           const float* conv_w, const float* conv_b, const float* gamma, const float* ffn_norm_w,
                           float* y, void* a_bf16);

int stts_test_ffn_fused(stts_engine* e, const void* a_bf16, const float* y, long long M, int C, const void* w1_bf16,
                        const float* b1, const void* w2_f16, const float* b2, const float* ffn_gamma, float* out,
                        void* out_bf16);
int stts_test_convnext_fused(stts_engine* e, const float* x, int B, int T, int C, const float* norm_w,
                             const float* conv_w, const float* conv_b, const float* gamma, const float* ffn_norm_w,
                             const void* w1_bf16, const float* b1, const void* w2_f16, const float* b2,
                             const float* ffn_gamma, float* out, void* out_bf16);


/* Chained DiT kernel (csrc/dit_chain.cuh): weights of the 12 blocks stacked per GEMM type, activation buffers of one
 * denoiser evaluation and up to 64 dependent phases (kind: 0 q|k|v|gate, 1 to_out, 2 w1|w3, 3 w2, 4 velocity, 5 joint
 * attention of block blk -- the attention fields below are only read when a phase of kind 5 is present).
 * `ready` must hold zeroed counters: n_phases per 128-row block rounded up to 32 ints, + 64. */
typedef struct stts_test_chain_args {
  const void *wqkvg, *wo, *w13, *w2, *wvel;                 /* bf16 */
  const float *bqkvg, *b13, *b2, *bvel, *qn, *kn, *cos_t, *sin_t;
  float* x; void* xb; float* stats; void* qkv; float* gate; void* ob; void* hb; float* vel; int32_t* ready;
  const int32_t* frames; const float* mod; const float* fold;
  uint64_t* trace; /* optional role timeline [CTAs][64][16] (tools/trace_chain.py), else NULL */
  const void *kv_ref, *kv_text;             /* cross K/V caches [12][2][B, R | P, 8, 128] bf16 */
  const int32_t *ref_len, *ph_len;          /* [B] */
  int32_t M, T, n_phases;
  int32_t B, R, P;
  int32_t qkv_db;                           /* 1: qkv is [2][3][M][1024], block i uses half i & 1 (required with kind 5) */
  int32_t kind[64];
  int32_t blk[64];
} stts_test_chain_args;
int stts_test_chain(stts_engine* e, const stts_test_chain_args* a);
/* fold table [stts_test_chain_fold_floats()] of the timestep whose adaLN table is a->mod, from the weights in `a` */
int stts_test_chain_fold(stts_engine* e, const stts_test_chain_args* a, float* fold_out);
int64_t stts_test_chain_fold_floats(void);
int stts_test_chain_stats_cast(stts_engine* e, const float* x, int M, const float* scale, void* xb, float* stats);

#ifdef __cplusplus
}
#endif
#endif /* SMALLTTS_B200_H_ */
