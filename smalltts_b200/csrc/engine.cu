// C-ABI implementation: weight registry + packing, condition encoder, DiT denoiser, DMD sampler, vocoder.
// Host orchestration only; all arithmetic is in gemm.cu (tcgen05) and kernels.cu.
#include <math.h>
#include <string.h>

#include <functional>
#include <array>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/smalltts_b200.h"
#include "dit_chain.cuh"
#include "gemm.cuh"
#include "kernels.cuh"
#include "launch.cuh"

using namespace stts;

namespace {

struct Err : std::runtime_error {
  int code;
  Err(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CK(expr)                                                                                         \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) {                                                                             \
      throw Err(STTS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" +  \
                                   std::to_string(__LINE__) + ")");                                     \
    }                                                                                                    \
  } while (0)

constexpr int D = 960, H = 8, HD = 120, HDP = 128, FF = 2400, NBLK = 12, LAT = 64;
constexpr int DP = 1024;  // 16 conv groups x (60 -> 64) padded channel layout of the conv position embedding
constexpr int MOD_LD = NBLK * 6 * D + 2 * D;  // per-timestep adaLN table: 12 x [6][960] + final [2][960]
constexpr int ROPE_MAX = 4096;                // dit.py:139, style.py:140, phonemes.py:196
const int VOC_R[6] = {8, 5, 5, 4, 2, 2};
const int VOC_C[7] = {2048, 1024, 512, 256, 128, 64, 32};
const int VOC_DEPTH[7] = {8, 3, 3, 3, 3, 3, 3};
const int ENC_R[6] = {2, 2, 4, 5, 5, 8};  // hf encoder config: downsampling_ratios
const int ENC_C[7] = {32, 64, 128, 256, 512, 1024, 2048};
const int ENC_DEPTH[7] = {3, 3, 3, 3, 3, 3, 8};

struct RawTensor {
  float* d = nullptr;
  std::vector<int64_t> shape;
  size_t numel = 0;
};

struct EncBlockW {
  bf16 *wqkvg, *wo, *w13, *w2;
  const float *qn, *kn, *an, *mn;
};
struct EncW {
  int d, heads, hd, inter, layers;
  float eps;
  std::vector<EncBlockW> blk;
  const float* final_norm;
};
struct DitBlockW {
  bf16 *wqkvg, *wo, *w13, *w2;
  float *bqkvg, *b13;
  const float *b2, *qn, *kn, *ada_w, *ada_b;
};
struct VocLayerW {
  const float *norm_w, *conv_w, *conv_b, *gamma, *ffn_norm_w, *b1, *b2, *ffn_gamma;
  bf16 *w1, *w2;
  void* w2h = nullptr;  // fp16 copy of 0.5 * linear2: the hidden activation is produced as 2*gelu in fp16
};

template <typename T>
struct Tmp {  // stream-ordered temporary
  T* p = nullptr;
  cudaStream_t st = nullptr;
  Tmp() {}
  Tmp(cudaStream_t s, size_t n) { alloc(s, n); }
  void alloc(cudaStream_t s, size_t n) {
    release();
    st = s;
    CK(cudaMallocAsync(reinterpret_cast<void**>(&p), (n ? n : 1) * sizeof(T), s));
  }
  void release() {
    if (p) cudaFreeAsync(p, st);
    p = nullptr;
  }
  ~Tmp() { release(); }
  Tmp(const Tmp&) = delete;
  Tmp& operator=(const Tmp&) = delete;
  operator T*() const { return p; }
};

}  // namespace

namespace {
struct Plan;
}

struct stts_cond {
  int B = 0, R = 0, P = 0;
  int *ref_len = nullptr, *ph_len = nullptr;  // device int32 [B]
  std::vector<int> h_ref_len, h_ph_len;
  bf16* kv_ref = nullptr;   // [12][2][B, R, 8, 128]
  bf16* kv_text = nullptr;  // [12][2][B, P, 8, 128]
  size_t ref_stride() const { return static_cast<size_t>(B) * R * H * HDP; }
  size_t text_stride() const { return static_cast<size_t>(B) * P * H * HDP; }
};

struct stts_engine {
  int device = 0;
  cudaStream_t st = nullptr;
  cudaStream_t st2 = nullptr;  // side stream: the style encoder runs beside the text encoder (fork/join on events)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::string err;
  std::map<std::string, RawTensor> raw[3];  // 0 = DiT, 1 = codec decoder (vocoder), 2 = codec encoder (optional)
  std::vector<void*> owned;  // packed buffers
  bool finalized = false, packed_once = false;
  bool owns_weights = true;  // false for stts_engine_clone handles: weight memory belongs to the source engine

  // packed DiT-side weights
  EncW style, text;
  bf16 *style_in_w, *style_out_w, *ph_proj_w, *wkv_ref, *wkv_text, *in_proj_w, *conv1_w, *conv2_w, *vel_w;
  float *style_in_b, *bkv_ref, *bkv_text, *in_proj_b, *conv1_b, *knorm_cross;
  bf16* in_proj_w_dense;
  const float* in_proj_b_dense;
  const float *style_out_b, *ph_proj_b, *conv2_b, *vel_b, *text_emb;
  std::vector<DitBlockW> blk;
  float *cos64, *sin64, *cos128, *sin128;
  std::map<uint32_t, float*> mod_cache;  // timestep bits -> device adaLN table [MOD_LD]
  // chained DiT path (dit_chain.cu): the 12 blocks' GEMM weights stacked per type + folded LayerNorm vectors per timestep
  ChainWeights chain_w;
  std::map<uint32_t, float*> fold_cache;  // timestep bits -> device fold table [kFoldFloats]
  // split-K scratch of the vocoder's wave-quantised GEMMs (gemm.cuh GemmShape::splits); engine stream only
  float* split_scratch = nullptr;
  int* split_cnt = nullptr;
  long long split_scratch_floats = 0, split_cnt_ints = 0;
  bool use_split = false;   // STTS_SPLITK=1: split-K for the wave-quantised vocoder GEMMs.  Off: measured slower (below)
  bool use_chain = true;    // STTS_NO_CHAIN=1: generic 8-launches-per-block path (also used for per-utterance timesteps)
  bool chain_split = false; // STTS_CHAIN_SPLIT=1: one GEMM per chain launch (debug / A-B timing of the in-kernel dependencies)
  // attention inside the chained kernel (csrc/chain_attn.cuh): 1 = a whole denoiser evaluation is ONE launch (61
  // phases), 2 = one launch per block (attention, to_out, w1|w3, w2, next q|k|v|gate), 0 = stand-alone attention kernel
  // between chain launches.  Default 0: measured on B200 at the headline shape the one-launch schedule is correct but
  // slower (3.72 vs 2.93 ms per 4-step loop): the item's serial chain (loads -> scores -> softmax ->
  // P V -> store -> publish) is longer than what the stand-alone kernel costs between two launches.
  int chain_attn = 0;       // STTS_CHAIN_ATTN=0|1|2

  // packed vocoder weights
  bf16* stem_w;
  const float* stem_b;
  std::vector<std::vector<VocLayerW>> voc;  // [7 stages][depth]
  bf16* up_w[6];
  float* up_b[6];
  const float *head_w, *head_b;

  // packed codec-encoder weights (model 2; only when its tensors were loaded)
  bool has_encoder = false, has_dit = false, has_decoder = false;
  std::vector<std::vector<VocLayerW>> enc;  // [7 stages][depth]
  bf16* enc_down_w[6];
  const float* enc_down_b[6];
  bf16* enc_head_w = nullptr;
  const float *enc_head_b = nullptr, *enc_stem_w = nullptr, *enc_stem_b = nullptr;

  // resampler filter banks per reduced (down, up) rate pair (built on first use, stts_resample)
  struct ResampleBank {
    int down, up, width, K;
    float* d;
  };
  std::map<std::pair<int, int>, ResampleBank> banks;

  // per-shape persistent plans (buffers + CUDA graphs) used by stts_synthesize
  std::map<std::array<int, 5>, Plan*> plans;
  uint64_t use_counter = 0;
  bool use_graphs = true;
  bool test_async = false;  // test hooks return without synchronising (micro-benchmarks time many launches)
  bool fused_ffn = true;   // STTS_NO_FUSED_FFN=1: C = 128 feed-forward as two GEMMs (debug / A-B timing)
  bool fused_tail = true;  // STTS_NO_FUSED_TAIL=1 falls back to mixer + two GEMMs for C <= 64 (debug / A-B timing)
  unsigned long long* seed_dev = nullptr;   // device u64 read by the Philox kernel
  unsigned long long* seed_host = nullptr;  // pinned staging for seed_dev

  stts_timing timing = {0, 0, 0, 0, 0};
  float voc_ms[2] = {0, 0};
  cudaEvent_t ev[8];
  cudaEvent_t ev_stop = nullptr;

  const RawTensor& W(int model, const std::string& name, std::initializer_list<int64_t> shape) {
    auto it = raw[model].find(name);
    if (it == raw[model].end()) throw Err(STTS_ERR_WEIGHTS, "missing weight: " + name);
    if (it->second.shape != std::vector<int64_t>(shape)) throw Err(STTS_ERR_WEIGHTS, "bad shape for weight: " + name);
    return it->second;
  }
  template <typename T>
  T* dalloc(size_t n, bool zero = true) {
    T* p = nullptr;
    CK(cudaMalloc(reinterpret_cast<void**>(&p), n * sizeof(T)));
    if (zero) CK(cudaMemsetAsync(p, 0, n * sizeof(T), st));
    owned.push_back(p);
    return p;
  }
};

namespace {

// ------------------------------------------------------------------ GEMM convenience wrappers
// Tile-width choice by a small cost model (cycles), from ncu measurements on B200:
//   * one SM pulls ~64 B/clk from L2, so a 64-deep k-iteration of a 128 x BN tile costs (16 KB + BN*128 B) / 64 =
//     256 + 2*BN cycles on its CTA (the MMA itself needs only 2*BN);
//   * the whole chip pulls ~6 KB/clk from L2, and narrow tiles re-read the A panel once per N tile, so the sum of all
//     tile loads is a second floor (this is what bounds the DiT GEMMs at M = 600);
//   * the GEMM is persistent (one CTA per SM, epilogue overlapped with the next tile).
int pick_bn(long long m, int n, int iters) {
  int best = 32;
  double best_cost = 1e30;
  for (int bn = 32; bn <= 256; bn *= 2) {
    if (bn > 32 && bn > n) break;
    const long long tiles = ((m + 127) / 128) * ((n + bn - 1) / bn);
    const long long rounds = (tiles + 147) / 148;
    const double per_cta = static_cast<double>(rounds) * iters * (256.0 + 2.0 * bn);
    const double chip = static_cast<double>(tiles) * iters * (16384.0 + 128.0 * bn) / 6000.0;
    const double cost = (per_cta > chip ? per_cta : chip) + 3000.0 + 30.0 * bn;
    if (cost < best_cost) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

// Split-K factor for a GEMM of `tiles` output tiles x `iters` k-iterations on a persistent grid of 148 CTAs: whole
// tiles come in rounds of 148, so 160 tiles take two rounds and 80 tiles of 128 iterations leave half of the chip
// idle; S parts per tile turn that into ceil(tiles S / 148) rounds of iters / S.  Only when it pays by >= 10 %, the
// parts keep >= 8 iterations and the fp32 partials fit the engine's scratch.
// MEASURED (tools/bench_splitk.py, profiles/r02_splitk_microbench.txt): it does not pay at these sizes -- the per-item
// cost of parking / re-reading 128 x BN fp32 partials and of the longer two-pass epilogue exceeds the k-iterations a
// part saves (600 x 8192 x 2048: 26.8 us whole tiles, 69.8 us with 2 parts; 600 x 2048 x 8192: 37.3 -> 44.6 / 48.0).
// The kernel path is kept (tested, opt-in with STTS_SPLITK=1); the engine computes whole tiles.
int pick_splits(const stts_engine* e, long long tiles, int iters, int bn) {
  if (!e->use_split || e->split_scratch == nullptr || (bn != 128 && bn != 256) || tiles >= 4 * 148) return 1;
  auto cost = [&](int S) {  // k-iterations on the critical path (+ the fix-up of a split tile)
    return static_cast<double>((tiles * S + 147) / 148) * ((iters + S - 1) / S) + (S > 1 ? 3.0 : 0.0);
  };
  int best = 1;
  for (int S = 2; S <= 4; ++S) {
    if (iters / S < 8) break;
    if (tiles * S * 128 * bn > e->split_scratch_floats || tiles * 8 > e->split_cnt_ints) break;
    if (cost(S) < cost(best)) best = S;
  }
  return cost(best) < 0.9 * cost(1) ? best : 1;
}

void set_splits(stts_engine* e, GemmShape& s, GemmEpi& epi, int bn) {
  const long long tiles = static_cast<long long>(s.B) * ((s.T + 127) / 128) * ((s.N + bn - 1) / bn) * s.groups;
  const bool epi_ok = epi.act == ACT_NONE || (epi.act == ACT_GELU && epi.gelu2_f16);
  s.splits = epi_ok ? pick_splits(e, tiles, s.taps * ((s.K + 63) / 64), bn) : 1;
  if (s.splits > 1) {
    epi.split_scratch = e->split_scratch;
    epi.split_counters = e->split_cnt;
  }
}

// Plain linear over flattened rows: out = epi(A[rows, K] * W[N, K]^T).  split_ok: the call is on the engine's main
// stream (the split-K scratch is per engine) and its epilogue has a split-K instantiation.
void linear(stts_engine* e, const bf16* a, long long rows, int k, int lda, const bf16* w, int n, int ldw, GemmEpi epi,
            int bn = 0, bool f16 = false, bool split_ok = false) {
  GemmShape s;
  s.ab_f16 = f16 ? 1 : 0;
  s.B = 1;
  s.T = static_cast<int>(rows);
  s.N = n;
  s.K = k;
  GemmA ga{a, k, lda};
  GemmW gw{w, n, ldw};
  if (bn == 0) bn = pick_bn(rows, n, (k + 63) / 64);
  if (split_ok) set_splits(e, s, epi, bn);
  CK(launch_gemm(e->st, bn, ga, gw, s, epi));
}

template <typename T>
const T* to_dev(stts_engine* e, const T* p, size_t n, int mem, Tmp<T>& hold) {
  if (mem == STTS_MEM_DEVICE) return p;
  hold.alloc(e->st, n);
  CK(cudaMemcpyAsync(hold.p, p, n * sizeof(T), cudaMemcpyHostToDevice, e->st));
  return hold.p;
}

void lens_to_dev(stts_engine* e, const int64_t* h, int n, int maxv, Tmp<int>& hold, std::vector<int>* keep = nullptr) {
  std::vector<int> v(n);
  for (int i = 0; i < n; ++i) {
    if (h[i] < 0 || h[i] > maxv) throw Err(STTS_ERR_INVALID, "length out of range");
    v[i] = static_cast<int>(h[i]);
  }
  hold.alloc(e->st, n);
  CK(cudaMemcpyAsync(hold.p, v.data(), n * sizeof(int), cudaMemcpyHostToDevice, e->st));
  CK(cudaStreamSynchronize(e->st));  // v goes out of scope
  if (keep) *keep = v;
}

// ------------------------------------------------------------------ weight packing
bf16* pack_lin(stts_engine* e, const RawTensor& t, int rows, int cols, bf16* dst = nullptr, int ld = 0, int row_off = 0,
               int row_mode = ROW_PLAIN, int col_mode = COL_PLAIN, float scale = 1.f) {
  if (ld == 0) ld = cols;
  if (dst == nullptr) dst = e->dalloc<bf16>(static_cast<size_t>(rows) * ld);
  CK(pack_matrix(e->st, t.d, rows, cols, scale, row_mode, row_off, col_mode, 0, dst, ld));
  return dst;
}

void build_encoder(stts_engine* e, EncW& w, const std::string& pre, int d, int heads, int inter, int layers, float eps) {
  w.d = d; w.heads = heads; w.hd = d / heads; w.inter = inter; w.layers = layers; w.eps = eps;
  for (int i = 0; i < layers; ++i) {
    const std::string p = pre + "blocks." + std::to_string(i) + ".";
    EncBlockW b;
    b.wqkvg = e->dalloc<bf16>(static_cast<size_t>(4) * d * d);
    const char* names[4] = {"wq", "wk", "wv", "gate"};
    for (int j = 0; j < 4; ++j) {
      pack_lin(e, e->W(0, p + "attention." + names[j] + ".weight", {d, d}), d, d, b.wqkvg, d, j * d);
    }
    b.wo = pack_lin(e, e->W(0, p + "attention.wo.weight", {d, d}), d, d);
    b.w13 = e->dalloc<bf16>(static_cast<size_t>(2) * inter * d);
    pack_lin(e, e->W(0, p + "mlp.w1.weight", {inter, d}), inter, d, b.w13, d, 0, ROW_INTERLEAVE16_LO);
    pack_lin(e, e->W(0, p + "mlp.w3.weight", {inter, d}), inter, d, b.w13, d, 0, ROW_INTERLEAVE16_HI);
    b.w2 = pack_lin(e, e->W(0, p + "mlp.w2.weight", {d, inter}), d, inter);
    b.qn = e->W(0, p + "attention.q_norm.weight", {heads, d / heads}).d;
    b.kn = e->W(0, p + "attention.k_norm.weight", {heads, d / heads}).d;
    b.an = e->W(0, p + "attention_norm.weight", {d}).d;
    b.mn = e->W(0, p + "mlp_norm.weight", {d}).d;
    w.blk.push_back(b);
  }
  w.final_norm = e->W(0, pre + "norm.weight", {d}).d;
}

// One ConvNeXt-1D layer's weights (hf:263-297) of codec model `model` (1 = decoder, 2 = encoder).
VocLayerW pack_convnext_layer(stts_engine* e, int model, const std::string& p, int c) {
  cudaStream_t st = e->st;
  VocLayerW w;
  w.gamma = e->W(model, p + "gamma", {c}).d;
  w.ffn_gamma = e->W(model, p + "ffn_gamma", {c}).d;
  w.norm_w = e->W(model, p + "norm.weight", {c}).d;
  w.ffn_norm_w = e->W(model, p + "ffn_norm.weight", {c}).d;
  w.w1 = pack_lin(e, e->W(model, p + "ffn.linear1.weight", {4 * c, c}), 4 * c, c);
  w.b1 = e->W(model, p + "ffn.linear1.bias", {4 * c}).d;
  w.w2 = nullptr;
  w.w2h = e->dalloc<uint16_t>(static_cast<size_t>(4) * c * c);
  CK(cast_f16(st, e->W(model, p + "ffn.linear2.weight", {c, 4 * c}).d, static_cast<long long>(4) * c * c, 0.5f, w.w2h));
  w.b2 = e->W(model, p + "ffn.linear2.bias", {c}).d;
  w.conv_w = e->W(model, p + "mixer.conv.weight", {c, 1, 7}).d;
  w.conv_b = e->W(model, p + "mixer.conv.bias", {c}).d;
  return w;
}

// Codec encoder (hf:300-403): mirror of the decoder with strided causal convolutions.
void finalize_encoder(stts_engine* e) {
  cudaStream_t st = e->st;
  e->enc_stem_w = e->W(2, "stem.conv.conv.weight", {32, 1, 7}).d;
  e->enc_stem_b = e->W(2, "stem.conv.conv.bias", {32}).d;
  e->enc.assign(7, {});
  for (int s = 0; s < 7; ++s) {
    const int c = ENC_C[s];
    for (int l = 0; l < ENC_DEPTH[s]; ++l) {
      const std::string p = (s == 0 ? std::string("stem.stage.") : "conv_layers." + std::to_string(s - 1) + ".stage.") +
                            std::to_string(l) + ".";
      e->enc[s].push_back(pack_convnext_layer(e, 2, p, c));
    }
    if (s < 6) {
      const int cout = ENC_C[s + 1], r = ENC_R[s];
      const std::string p = "conv_layers." + std::to_string(s) + ".conv.conv.";
      e->enc_down_w[s] = e->dalloc<bf16>(static_cast<size_t>(cout) * 2 * r * c);
      CK(pack_conv_strided(st, e->W(2, p + "weight", {cout, c, 2 * r}).d, cout, c, r, e->enc_down_w[s]));
      e->enc_down_b[s] = e->W(2, p + "bias", {cout}).d;
    }
  }
  e->enc_head_w = e->dalloc<bf16>(static_cast<size_t>(LAT) * 7 * 2048);
  CK(pack_conv_taps(st, e->W(2, "head.conv.weight", {LAT, 2048, 7}).d, LAT, 2048, 7, 2048, LAT, LAT, e->enc_head_w, 7 * 2048));
  e->enc_head_b = e->W(2, "head.conv.bias", {LAT}).d;
  e->has_encoder = true;
}

void build_rope(stts_engine* e, int rot, float** cos_d, float** sin_d) {
  // angle[p, i] = p * 10000^(-2i/rot) in fp32 like dit.py:141-143 / style.py:13-16; cos/sin of that fp32 angle
  const int half = rot / 2;
  std::vector<float> c(static_cast<size_t>(ROPE_MAX) * half), s(c.size());
  for (int i = 0; i < half; ++i) {
    const float inv = 1.0f / powf(10000.0f, static_cast<float>(2 * i) / static_cast<float>(rot));
    for (int p = 0; p < ROPE_MAX; ++p) {
      const float ang = static_cast<float>(p) * inv;
      c[static_cast<size_t>(p) * half + i] = static_cast<float>(cos(static_cast<double>(ang)));
      s[static_cast<size_t>(p) * half + i] = static_cast<float>(sin(static_cast<double>(ang)));
    }
  }
  *cos_d = e->dalloc<float>(c.size(), false);
  *sin_d = e->dalloc<float>(s.size(), false);
  CK(cudaMemcpyAsync(*cos_d, c.data(), c.size() * 4, cudaMemcpyHostToDevice, e->st));
  CK(cudaMemcpyAsync(*sin_d, s.data(), s.size() * 4, cudaMemcpyHostToDevice, e->st));
  CK(cudaStreamSynchronize(e->st));
}

void finalize_dit(stts_engine* e) {
  cudaStream_t st = e->st;
  // ---- style encoder (style.py:118-174).  exp(log_scale) (style.py:167) is folded into in_proj.
  float log_scale = 0.f;
  {
    auto it = e->raw[0].find("style_encoder.log_scale");
    if (it == e->raw[0].end() || it->second.numel != 1) throw Err(STTS_ERR_WEIGHTS, "missing/bad weight: style_encoder.log_scale");
    CK(cudaMemcpy(&log_scale, it->second.d, 4, cudaMemcpyDeviceToHost));
  }
  const float sscale = expf(log_scale);
  e->style_in_w = pack_lin(e, e->W(0, "style_encoder.in_proj.weight", {512, 64}), 512, 64, nullptr, 0, 0, ROW_PLAIN,
                           COL_PLAIN, sscale);
  e->style_in_b = e->dalloc<float>(512);
  CK(pack_vector(st, e->W(0, "style_encoder.in_proj.bias", {512}).d, 512, sscale, ROW_PLAIN, 0, e->style_in_b));
  build_encoder(e, e->style, "style_encoder.", 512, 8, 1536, 12, 1e-5f);
  e->style_out_w = pack_lin(e, e->W(0, "style_encoder.out_proj.weight", {D, 512}), D, 512);
  e->style_out_b = e->W(0, "style_encoder.out_proj.bias", {D}).d;
  // ---- text encoder (phonemes.py:170-207)
  e->text_emb = e->W(0, "phoneme_embedding.text_embedding.weight", {198, 512}).d;
  build_encoder(e, e->text, "phoneme_embedding.", 512, 4, 1024, 8, 1e-6f);
  e->ph_proj_w = pack_lin(e, e->W(0, "dit.phoneme_proj.weight", {D, 512}), D, 512);
  e->ph_proj_b = e->W(0, "dit.phoneme_proj.bias", {D}).d;
  // ---- DiT input embedding (dit.py:215-253)
  // The grouped k=31 convs read 16 groups of 60 channels.  TMA box starts must be 16-byte aligned, so the conv
  // input lives in a padded layout (group g at columns [64g, 64g+60), pads zero): proj emits that layout directly.
  e->in_proj_w = e->dalloc<bf16>(static_cast<size_t>(DP) * 64);
  CK(pack_conv_taps(st, e->W(0, "dit.input_embed.proj.weight", {D, 64}).d, D, 64, 1, 64, 60, 64, e->in_proj_w, 64));
  e->in_proj_b = e->dalloc<float>(DP);
  CK(pack_vector(st, e->W(0, "dit.input_embed.proj.bias", {D}).d, D, 1.f, ROW_GROUPPAD_60_64, 0, e->in_proj_b));
  // ... and once more densely for the fp32 residual h (dit.py:252: conv_pos_embed(x) + x)
  e->in_proj_w_dense = pack_lin(e, e->W(0, "dit.input_embed.proj.weight", {D, 64}), D, 64);
  e->in_proj_b_dense = e->W(0, "dit.input_embed.proj.bias", {D}).d;
  for (int c = 0; c < 2; ++c) {
    const std::string p = std::string("dit.input_embed.conv_pos_embed.conv") + (c ? "2" : "1");
    bf16* w;
    if (c == 0) {  // conv1: 16 groups, padded 64-channel input AND output layout
      w = e->dalloc<bf16>(static_cast<size_t>(16) * 64 * 31 * 64);
      CK(pack_conv_taps(st, e->W(0, p + ".weight", {D, 60, 31}).d, D, 60, 31, 64, 60, 64, w, 31 * 64));
    } else {  // conv2: dense [M, 960] output in 15 tiles of 64 columns, each over two adjacent padded input groups
      w = e->dalloc<bf16>(static_cast<size_t>(D) * 31 * 128);
      CK(pack_conv_dense_tiles(st, e->W(0, p + ".weight", {D, 60, 31}).d, 31, w));
    }
    (c ? e->conv2_w : e->conv1_w) = w;
    if (c == 0) {  // conv1 writes the padded layout again (pad columns: zero weights + zero bias -> mish(0) = 0)
      e->conv1_b = e->dalloc<float>(DP);
      CK(pack_vector(st, e->W(0, p + ".bias", {D}).d, D, 1.f, ROW_GROUPPAD_60_64, 0, e->conv1_b));
    } else {
      e->conv2_b = e->W(0, p + ".bias", {D}).d;
    }
  }
  // ---- time / adaLN path stays fp32 (GEMV): only check presence
  e->W(0, "time_embedding.mlp.0.weight", {D, 256}); e->W(0, "time_embedding.mlp.0.bias", {D});
  e->W(0, "time_embedding.mlp.2.weight", {D, D});   e->W(0, "time_embedding.mlp.2.bias", {D});
  e->W(0, "dit.emb_proj.0.weight", {2 * D, D});     e->W(0, "dit.emb_proj.0.bias", {2 * D});
  e->W(0, "dit.emb_proj.2.weight", {D, 2 * D});     e->W(0, "dit.emb_proj.2.bias", {D});
  e->W(0, "dit.norm_out.linear.weight", {2 * D, D}); e->W(0, "dit.norm_out.linear.bias", {2 * D});
  // ---- blocks + fused cross-KV projection for all 12 blocks (dit.py:80-93)
  e->wkv_ref = e->dalloc<bf16>(static_cast<size_t>(NBLK) * 2 * D * D);
  e->wkv_text = e->dalloc<bf16>(static_cast<size_t>(NBLK) * 2 * D * D);
  e->bkv_ref = e->dalloc<float>(NBLK * 2 * D);
  e->bkv_text = e->dalloc<float>(NBLK * 2 * D);
  e->knorm_cross = e->dalloc<float>(NBLK * H * HD);
  // GEMM weights of the 12 blocks live stacked per type (one TMA tensor map per type in the chained kernel); the
  // per-block pointers of the generic path point into the same memory.
  bf16* wo_all = e->dalloc<bf16>(static_cast<size_t>(NBLK) * D * H * HDP);
  bf16* w13_all = e->dalloc<bf16>(static_cast<size_t>(NBLK) * 2 * FF * D);
  bf16* w2_all = e->dalloc<bf16>(static_cast<size_t>(NBLK) * D * FF);
  bf16* wqkvg_pad = e->dalloc<bf16>(static_cast<size_t>(NBLK) * kChainQKVG * D);  // q|k|v head-padded 120 -> 128, then gate
  float* bqkvg_pad = e->dalloc<float>(static_cast<size_t>(NBLK) * kChainQKVG);
  float* b13_all = e->dalloc<float>(static_cast<size_t>(NBLK) * 2 * FF);
  float* b2_all = e->dalloc<float>(static_cast<size_t>(NBLK) * D);
  float* qn_all = e->dalloc<float>(static_cast<size_t>(NBLK) * H * HD);
  float* kn_all = e->dalloc<float>(static_cast<size_t>(NBLK) * H * HD);
  for (int i = 0; i < NBLK; ++i) {
    const std::string p = "dit.transformer_blocks." + std::to_string(i) + ".";
    DitBlockW b;
    b.wqkvg = e->dalloc<bf16>(static_cast<size_t>(4) * D * D);
    b.bqkvg = e->dalloc<float>(4 * D);
    const char* names[4] = {"to_q", "to_k_self", "to_v_self", "gate"};
    bf16* wq_pad = wqkvg_pad + static_cast<size_t>(i) * kChainQKVG * D;
    float* bq_pad = bqkvg_pad + static_cast<size_t>(i) * kChainQKVG;
    for (int j = 0; j < 4; ++j) {
      const RawTensor& wj = e->W(0, p + "attn." + names[j] + ".weight", {D, D});
      pack_lin(e, wj, D, D, b.wqkvg, D, j * D);
      if (j < 3) {
        const float* bj = e->W(0, p + "attn." + names[j] + ".bias", {D}).d;
        CK(pack_vector(st, bj, D, 1.f, ROW_PLAIN, j * D, b.bqkvg));
        CK(pack_rows_headpad(st, wj.d, D, D, j * H * HDP, wq_pad, D));
        CK(pack_vec_headpad(st, bj, D, j * H * HDP, bq_pad));
      } else {
        pack_lin(e, wj, D, D, wq_pad, D, 3 * H * HDP);
      }
    }
    // to_out consumes head-padded (120 -> 128) attention output: K = 8 * 128
    b.wo = wo_all + static_cast<size_t>(i) * D * H * HDP;
    pack_lin(e, e->W(0, p + "attn.to_out.0.weight", {D, D}), D, D, b.wo, H * HDP, 0, ROW_PLAIN, COL_HEADPAD_120_128);
    b.w13 = w13_all + static_cast<size_t>(i) * 2 * FF * D;
    b.b13 = b13_all + static_cast<size_t>(i) * 2 * FF;
    pack_lin(e, e->W(0, p + "ff.w1.weight", {FF, D}), FF, D, b.w13, D, 0, ROW_INTERLEAVE16_LO);
    pack_lin(e, e->W(0, p + "ff.w3.weight", {FF, D}), FF, D, b.w13, D, 0, ROW_INTERLEAVE16_HI);
    CK(pack_vector(st, e->W(0, p + "ff.w1.bias", {FF}).d, FF, 1.f, ROW_INTERLEAVE16_LO, 0, b.b13));
    CK(pack_vector(st, e->W(0, p + "ff.w3.bias", {FF}).d, FF, 1.f, ROW_INTERLEAVE16_HI, 0, b.b13));
    b.w2 = pack_lin(e, e->W(0, p + "ff.w2.weight", {D, FF}), D, FF, w2_all + static_cast<size_t>(i) * D * FF);
    b.b2 = e->W(0, p + "ff.w2.bias", {D}).d;
    b.qn = e->W(0, p + "attn.q_norm.weight", {H, HD}).d;
    b.kn = e->W(0, p + "attn.k_norm.weight", {H, HD}).d;
    CK(cudaMemcpyAsync(b2_all + static_cast<size_t>(i) * D, b.b2, D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(qn_all + static_cast<size_t>(i) * H * HD, b.qn, H * HD * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(kn_all + static_cast<size_t>(i) * H * HD, b.kn, H * HD * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(e->knorm_cross + static_cast<size_t>(i) * H * HD, e->W(0, p + "attn.k_norm_cross.weight", {H, HD}).d,
                       H * HD * sizeof(float), cudaMemcpyDeviceToDevice, st));
    b.ada_w = e->W(0, p + "attn_norm.linear.weight", {6 * D, D}).d;
    b.ada_b = e->W(0, p + "attn_norm.linear.bias", {6 * D}).d;
    e->blk.push_back(b);
    const char* kv[4] = {"to_k_ref", "to_v_ref", "to_k_text", "to_v_text"};
    for (int j = 0; j < 4; ++j) {
      bf16* dw = j < 2 ? e->wkv_ref : e->wkv_text;
      float* db = j < 2 ? e->bkv_ref : e->bkv_text;
      const int row_off = i * 2 * D + (j & 1) * D;
      pack_lin(e, e->W(0, p + "attn." + kv[j] + ".weight", {D, D}), D, D, dw, D, row_off);
      CK(pack_vector(st, e->W(0, p + "attn." + kv[j] + ".bias", {D}).d, D, 1.f, ROW_PLAIN, row_off, db));
    }
  }
  e->vel_w = pack_lin(e, e->W(0, "velocity.weight", {LAT, D}), LAT, D);
  e->vel_b = e->W(0, "velocity.bias", {LAT}).d;
  build_rope(e, 64, &e->cos64, &e->sin64);
  build_rope(e, 128, &e->cos128, &e->sin128);
  e->chain_w.wqkvg = wqkvg_pad; e->chain_w.wo = wo_all; e->chain_w.w13 = w13_all; e->chain_w.w2 = w2_all;
  e->chain_w.wvel = e->vel_w; e->chain_w.bqkvg = bqkvg_pad; e->chain_w.b13 = b13_all; e->chain_w.b2 = b2_all;
  e->chain_w.bvel = e->vel_b; e->chain_w.qn = qn_all; e->chain_w.kn = kn_all;
  e->chain_w.cos_t = e->cos64; e->chain_w.sin_t = e->sin64;
  e->has_dit = true;
}

void finalize_decoder(stts_engine* e) {
  cudaStream_t st = e->st;
  // ---- vocoder (hf:406-500)
  e->stem_w = e->dalloc<bf16>(static_cast<size_t>(2048) * 7 * 64);
  CK(pack_conv_taps(st, e->W(1, "stem.conv.conv.weight", {2048, 64, 7}).d, 2048, 64, 7, 64, 2048, 2048, e->stem_w, 7 * 64));
  e->stem_b = e->W(1, "stem.conv.conv.bias", {2048}).d;
  e->voc.resize(7);
  for (int s = 0; s < 7; ++s) {
    const int c = VOC_C[s];
    for (int l = 0; l < VOC_DEPTH[s]; ++l) {
      const std::string p = (s == 0 ? std::string("stem.stage.") : "conv_layers." + std::to_string(s - 1) + ".stage.") +
                            std::to_string(l) + ".";
      VocLayerW w = pack_convnext_layer(e, 1, p, c);
      e->voc[s].push_back(w);
    }
    if (s < 6) {
      const int cin = VOC_C[s], cout = VOC_C[s + 1], r = VOC_R[s];
      const std::string p = "conv_layers." + std::to_string(s) + ".convtr.convtr.";
      e->up_w[s] = e->dalloc<bf16>(static_cast<size_t>(r) * cout * 2 * cin);
      CK(pack_convtr(st, e->W(1, p + "weight", {cin, cout, 2 * r}).d, cin, cout, r, e->up_w[s]));
      e->up_b[s] = e->dalloc<float>(r * cout);
      CK(tile_vector(st, e->W(1, p + "bias", {cout}).d, cout, r, e->up_b[s]));
    }
  }
  e->head_w = e->W(1, "head.conv.weight", {1, 32, 7}).d;
  e->head_b = e->W(1, "head.conv.bias", {1}).d;
  if (e->split_scratch == nullptr) {  // fp32 partial tiles of the split-K GEMMs (largest user: 160 tiles x 3 parts x 128 x 256)
    e->split_scratch_floats = 24ll << 20;  // 96 MB
    e->split_cnt_ints = 64 << 10;
    e->split_scratch = e->dalloc<float>(static_cast<size_t>(e->split_scratch_floats), false);
    e->split_cnt = e->dalloc<int>(static_cast<size_t>(e->split_cnt_ints), true);
  }
  e->has_decoder = true;
}

// Each of the three models is packed iff any of its tensors was loaded (and then all of them must be there): a
// codec-only engine serves the reference's standalone Decoder / Encoder (codec/onnx.py:34-75, used alone by sv.py:24).
void finalize(stts_engine* e) {
  if (e->raw[0].empty() && e->raw[1].empty() && e->raw[2].empty()) throw Err(STTS_ERR_WEIGHTS, "no weights were loaded");
  if (!e->raw[0].empty()) finalize_dit(e);
  if (!e->raw[1].empty()) finalize_decoder(e);
  if (!e->raw[2].empty()) finalize_encoder(e);
  CK(cudaStreamSynchronize(e->st));
  e->finalized = true;
}

// ------------------------------------------------------------------ condition encoder
// One pre-norm transformer block of style.py:74-105 / phonemes.py:136-167 on x [B*N, d] (fp32, in place).
void encoder_block(stts_engine* e, const EncW& W, const EncBlockW& bw, float* x, int B, int N, const int* len_dev) {
  cudaStream_t st = e->st;
  const long long M = static_cast<long long>(B) * N;
  const int d = W.d;
  Tmp<bf16> a(st, M * d), qb(st, M * d), kb(st, M * d), vb(st, M * d), ob(st, M * d), hb(st, M * W.inter);
  Tmp<float> qkvg(st, M * 4 * d);
  CK(rms_norm_bf16(st, x, M, d, bw.an, W.eps, a));
  GemmEpi ep;
  ep.out_f32 = qkvg; ep.ld_out = 4 * d;
  linear(e, a, M, d, d, bw.wqkvg, 4 * d, d, ep);
  const float* cs = W.hd == 64 ? e->cos64 : e->cos128;
  const float* sn = W.hd == 64 ? e->sin64 : e->sin128;
  CK(head_split_qkv_bf16(st, qkvg, 4 * d, d, M, N, W.heads, W.hd, W.hd, bw.qn, bw.kn, W.eps, W.hd, cs, sn, qb, kb, vb));
  AttnSeg seg;
  seg.k = kb; seg.v = vb; seg.len = len_dev; seg.n_max = N;
  CK(attention_bf16(st, qb, B, N, W.heads, W.hd, W.hd, &seg, 1, qkvg, 4 * d, 3 * d, ob));
  GemmEpi eo;
  eo.residual = x; eo.ld_res = d; eo.out_f32 = x; eo.ld_out = d;
  linear(e, ob, M, d, d, bw.wo, d, d, eo);
  CK(rms_norm_bf16(st, x, M, d, bw.mn, W.eps, a));
  GemmEpi e1;
  e1.act = ACT_SWIGLU16; e1.out_bf16 = hb; e1.ld_out = W.inter;
  linear(e, a, M, d, d, bw.w13, 2 * W.inter, d, e1);
  GemmEpi e2;
  e2.residual = x; e2.ld_res = d; e2.out_f32 = x; e2.ld_out = d;
  linear(e, hb, M, W.inter, W.inter, bw.w2, d, W.inter, e2);
}

void encode_conditions_core(stts_engine* e, stts_cond* c, const float* dref, const long long* dids);

stts_cond* encode_conditions(stts_engine* e, const float* ref, const int64_t* ref_len, const int64_t* ids,
                             const int64_t* ph_len, int B, int R, int P, int mem) {
  if (B < 1 || R < 1 || P < 1 || R > ROPE_MAX || P > ROPE_MAX) throw Err(STTS_ERR_INVALID, "bad B/R/P");
  cudaStream_t st = e->st;
  stts_cond* c = new stts_cond();
  try {
    c->B = B; c->R = R; c->P = P;
    CK(cudaMalloc(reinterpret_cast<void**>(&c->ref_len), B * sizeof(int)));
    CK(cudaMalloc(reinterpret_cast<void**>(&c->ph_len), B * sizeof(int)));
    {
      Tmp<int> t1, t2;
      lens_to_dev(e, ref_len, B, R, t1, &c->h_ref_len);
      lens_to_dev(e, ph_len, B, P, t2, &c->h_ph_len);
      CK(cudaMemcpyAsync(c->ref_len, t1.p, B * sizeof(int), cudaMemcpyDeviceToDevice, st));
      CK(cudaMemcpyAsync(c->ph_len, t2.p, B * sizeof(int), cudaMemcpyDeviceToDevice, st));
    }
    CK(cudaMalloc(reinterpret_cast<void**>(&c->kv_ref), NBLK * 2 * c->ref_stride() * sizeof(bf16)));
    CK(cudaMalloc(reinterpret_cast<void**>(&c->kv_text), NBLK * 2 * c->text_stride() * sizeof(bf16)));

    Tmp<float> href;
    Tmp<long long> hids;
    const float* dref = to_dev(e, ref, static_cast<size_t>(B) * R * LAT, mem, href);
    const long long* dids = reinterpret_cast<const long long*>(
        to_dev(e, reinterpret_cast<const long long*>(ids), static_cast<size_t>(B) * P, mem, hids));
    encode_conditions_core(e, c, dref, dids);
    CK(cudaStreamSynchronize(st));
  } catch (...) {
    cudaFree(c->ref_len); cudaFree(c->ph_len); cudaFree(c->kv_ref); cudaFree(c->kv_text);
    delete c;
    throw;
  }
  return c;
}

// Launch-only part of the condition encoder (no allocation of persistent state, no synchronisation): safe to
// capture into a CUDA graph.  `c` carries device lengths and the K/V cache buffers to fill.
void encode_conditions_core(stts_engine* e, stts_cond* c, const float* dref, const long long* dids) {
  cudaStream_t main_st = e->st;
  const int B = c->B, R = c->R, P = c->P;
  // The style path (src 0, B*R rows) and the text path (src 1, B*P rows) are independent chains of small kernels:
  // fork the style path onto the side stream and join before returning (works eagerly and under stream capture).
  struct Restore {
    stts_engine* e;
    cudaStream_t s;
    ~Restore() { e->st = s; }
  } restore{e, main_st};
  CK(cudaEventRecord(e->ev_fork, main_st));
  CK(cudaStreamWaitEvent(e->st2, e->ev_fork, 0));
  {
    for (int src = 0; src < 2; ++src) {
      e->st = src == 0 ? e->st2 : main_st;
      cudaStream_t st = e->st;
      const int N = src == 0 ? R : P;
      const long long M = static_cast<long long>(B) * N;
      const int* len_dev = src == 0 ? c->ref_len : c->ph_len;
      const EncW& W = src == 0 ? e->style : e->text;
      Tmp<float> x(st, M * 512);
      Tmp<bf16> a(st, M * 512), seq(st, M * D);
      if (src == 0) {  // style.py:166-167
        Tmp<bf16> rb(st, M * LAT);
        CK(cast_bf16(st, dref, M * LAT, rb));
        GemmEpi ep;
        ep.bias = e->style_in_b; ep.out_f32 = x; ep.ld_out = 512;
        linear(e, rb, M, LAT, LAT, e->style_in_w, 512, LAT, ep);
      } else {  // phonemes.py:203
        CK(embed_gather(st, dids, M, e->text_emb, 198, 512, x));
      }
      for (int i = 0; i < W.layers; ++i) encoder_block(e, W, W.blk[i], x, B, N, len_dev);
      CK(rms_norm_bf16(st, x, M, 512, W.final_norm, W.eps, a));
      // out_proj / phoneme_proj with masked rows zeroed (style.py:172-173 / dit.py:293-298)
      GemmEpi ep;
      ep.bias = src == 0 ? e->style_out_b : e->ph_proj_b;
      ep.row_len = len_dev; ep.rows_per_batch = N;
      ep.out_bf16 = seq; ep.ld_out = D;
      linear(e, a, M, 512, 512, src == 0 ? e->style_out_w : e->ph_proj_w, D, 512, ep);
      // all 12 blocks' cross K/V in one GEMM, then per-head k_norm_cross (dit.py:80-93)
      const int NKV = NBLK * 2 * D;
      Tmp<float> kvf(st, M * NKV);
      GemmEpi ek;
      ek.bias = src == 0 ? e->bkv_ref : e->bkv_text; ek.out_f32 = kvf; ek.ld_out = NKV;
      linear(e, seq, M, D, D, src == 0 ? e->wkv_ref : e->wkv_text, NKV, D, ek);
      bf16* cache = src == 0 ? c->kv_ref : c->kv_text;
      const size_t stride = src == 0 ? c->ref_stride() : c->text_stride();
      CK(kv_split_bf16(st, kvf, NKV, M, NBLK, H, HD, HDP, 1e-6f, e->knorm_cross, cache, static_cast<long long>(stride)));
    }
  }
  e->st = main_st;
  CK(cudaEventRecord(e->ev_join, e->st2));
  CK(cudaStreamWaitEvent(main_st, e->ev_join, 0));
}

// ------------------------------------------------------------------ adaLN tables (function of t only)
// model.py:23-30 -> dit.py:270-274 -> per block dit.py:19-23 (gates stored as tanh) -> final dit.py:36-37
void compute_mod(stts_engine* e, const float* t_dev, int rows, float* table /*[rows][MOD_LD]*/) {
  cudaStream_t st = e->st;
  auto& R = e->raw[0];
  Tmp<float> f(st, rows * 256), h(st, rows * D), te(st, rows * D), h2(st, rows * 2 * D), emb(st, rows * D);
  CK(time_features(st, t_dev, rows, f));
  CK(gemv_rows(st, f, rows, 256, R["time_embedding.mlp.0.weight"].d, R["time_embedding.mlp.0.bias"].d, D, 0, 1, 0, 0, h, D));
  CK(gemv_rows(st, h, rows, D, R["time_embedding.mlp.2.weight"].d, R["time_embedding.mlp.2.bias"].d, D, 0, 0, 0, 0, te, D));
  CK(gemv_rows(st, te, rows, D, R["dit.emb_proj.0.weight"].d, R["dit.emb_proj.0.bias"].d, 2 * D, 0, 1, 0, 0, h2, 2 * D));
  CK(gemv_rows(st, h2, rows, 2 * D, R["dit.emb_proj.2.weight"].d, R["dit.emb_proj.2.bias"].d, D, 0, 0, 0, 0, emb, D));
  for (int i = 0; i < NBLK; ++i) {
    // chunks: shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp -> tanh on chunks 2 and 5
    CK(gemv_rows(st, emb, rows, D, e->blk[i].ada_w, e->blk[i].ada_b, 6 * D, 1, 0, D, 0x24u, table + i * 6 * D, MOD_LD));
  }
  CK(gemv_rows(st, emb, rows, D, R["dit.norm_out.linear.weight"].d, R["dit.norm_out.linear.bias"].d, 2 * D, 1, 0, 0, 0,
               table + NBLK * 6 * D, MOD_LD));
}

const float* cached_mod(stts_engine* e, float t) {
  uint32_t bits;
  memcpy(&bits, &t, 4);
  auto it = e->mod_cache.find(bits);
  if (it != e->mod_cache.end()) return it->second;
  float* table = nullptr;
  CK(cudaMalloc(reinterpret_cast<void**>(&table), MOD_LD * sizeof(float)));
  e->owned.push_back(table);
  Tmp<float> td(e->st, 1);
  CK(cudaMemcpyAsync(td.p, &t, 4, cudaMemcpyHostToDevice, e->st));
  CK(cudaStreamSynchronize(e->st));
  compute_mod(e, td, 1, table);
  e->mod_cache[bits] = table;
  return table;
}

// Folded LayerNorm vectors of the chained DiT path for timestep t (dit_chain.cuh); nullptr when that path is off.
const float* cached_fold(stts_engine* e, float t) {
  if (!e->use_chain) return nullptr;
  uint32_t bits;
  memcpy(&bits, &t, 4);
  auto it = e->fold_cache.find(bits);
  if (it != e->fold_cache.end()) return it->second;
  const float* mod = cached_mod(e, t);
  float* fold = nullptr;
  CK(cudaMalloc(reinterpret_cast<void**>(&fold), static_cast<size_t>(kFoldFloats) * sizeof(float)));
  e->owned.push_back(fold);
  CK(chain_fold_table(e->st, e->chain_w, mod, fold));
  CK(cudaStreamSynchronize(e->st));
  e->fold_cache[bits] = fold;
  return fold;
}

// ------------------------------------------------------------------ denoiser (dit.py:316-327, model.py:97-100)
constexpr int kChainLaunches = NBLK + 1;  // q|k|v|gate of block 0, then one launch per block (the last ends in the velocity head)
constexpr int kEvalPhases = 1 + 5 * NBLK;  // one-launch evaluation: q|k|v|gate, then (attention, to_out, w1|w3, w2, next) x 12
static_assert(kEvalPhases <= kChainMaxPhases, "phase table of the chained kernel too small");
struct DenoiseWs {
  Tmp<float> h, x, qkvg, v, stats;
  Tmp<bf16> hm, c1, a, qkv, ob, hb;
  Tmp<int> ready;
  bf16 *qb = nullptr, *kb = nullptr, *vb = nullptr;  // views of qkv: [3][M][8][128]
  void alloc(cudaStream_t st, long long M) {
    h.alloc(st, M * D); x.alloc(st, M * D); qkvg.alloc(st, M * 4 * D);
    hm.alloc(st, M * DP); c1.alloc(st, M * DP); a.alloc(st, M * D);
    // two q|k|v buffers (by block parity) for the launches that run attention inside the chain.  Zeroed once: attention
    // fetches whole 16-key boxes, and a masked key's V row must never hold a NaN pattern left behind by an earlier owner
    // of the memory (0 x NaN), even before its row block has been written for the first time.
    qkv.alloc(st, 2 * 3 * M * H * HDP); ob.alloc(st, M * H * HDP);
    CK(cudaMemsetAsync(qkv.p, 0, static_cast<size_t>(2) * 3 * M * H * HDP * sizeof(bf16), st));
    qb = qkv.p; kb = qb + M * H * HDP; vb = kb + M * H * HDP;
    hb.alloc(st, M * FF);
    stats.alloc(st, M * 2 * kChainParts);
    ready.alloc(st, ready_ints(M));
  }
  static size_t ready_ints(long long M) {  // the largest of the three launch schedules of denoise()
    const size_t a = static_cast<size_t>(kChainLaunches) * chain_ready_ints(static_cast<int>(M), 5);
    const size_t b = chain_ready_ints(static_cast<int>(M), kEvalPhases);
    return a > b ? a : b;
  }
};

// `fold` (cached_fold of the same timestep, shared by all rows) selects the chained path: per block one attention launch
// and one persistent GEMM-chain launch (dit_chain.cu) instead of eight kernels.
void denoise(stts_engine* e, const stts_cond* c, DenoiseWs& ws, const bf16* xt_bf16, const int* frames_dev,
             const float* mod, int ld_mod, const float* fold, int B, int T, float* v_out) {
  cudaStream_t st = e->st;
  const long long M = static_cast<long long>(B) * T;
  // input embedding: h = proj(x); hm = masked bf16 copy; x = mask(mish(conv2(mask(mish(conv1(hm)))))) + h
  {
    GemmEpi ep;  // masked bf16 copy in the group-padded layout (conv input)
    ep.bias = e->in_proj_b; ep.row_len = frames_dev; ep.rows_per_batch = T;
    ep.out_bf16 = ws.hm; ep.ld_out = DP;
    linear(e, xt_bf16, M, LAT, LAT, e->in_proj_w, DP, LAT, ep);
    GemmEpi eh;  // dense fp32 h (unmasked) for the residual
    eh.bias = e->in_proj_b_dense; eh.out_f32 = ws.h; eh.ld_out = D;
    linear(e, xt_bf16, M, LAT, LAT, e->in_proj_w_dense, D, LAT, eh);
    GemmShape s;
    s.B = B; s.T = T; s.N = 64; s.K = 64; s.taps = 31; s.tap_shift0 = -15; s.tap_step = 1;
    s.groups = 16; s.a_group_koff = 64; s.w_group_rows = 64; s.out_group_cols = 64; s.res_group_cols = 64;
    GemmEpi e1;
    e1.bias = e->conv1_b; e1.act = ACT_MISH; e1.row_len = frames_dev; e1.out_bf16 = ws.c1; e1.ld_out = DP;
    CK(launch_gemm(st, 64, GemmA{ws.hm, DP, DP}, GemmW{e->conv1_w, 16 * 64, 31 * 64}, s, e1));
    // conv2: 15 dense 64-column tiles; tile n reads padded groups n and n+1 (K = 128)
    s.groups = 15; s.K = 128;
    GemmEpi e2;
    e2.bias = e->conv2_b; e2.act = ACT_MISH; e2.row_len = frames_dev; e2.residual = ws.h; e2.ld_res = D;
    e2.out_f32 = ws.x; e2.ld_out = D;
    CK(launch_gemm(st, 64, GemmA{ws.c1, DP, DP}, GemmW{e->conv2_w, D, 31 * 128}, s, e2));
  }
  if (fold != nullptr && ld_mod == 0 && e->use_chain) {
    const int Mi = static_cast<int>(M);
    // attention inside the chain: all keys of an utterance in one 256-column accumulator, and the per-row-block target
    // table must fit (M <= 64 x 128 rows); other shapes run the stand-alone attention kernel between chain launches
    const int attn_mode =
        (Mi + 127) / 128 <= kChainMaxRowBlocks && chain_attn_fits(T, c->R, c->P) && !e->chain_split ? e->chain_attn : 0;
    const int ready_ints = chain_ready_ints(Mi, 5);
    ChainBuffers cb;
    cb.x = ws.x; cb.xb = ws.a; cb.stats = ws.stats; cb.qkv = ws.qkv; cb.gate = ws.qkvg; cb.ob = ws.ob; cb.hb = ws.hb;
    cb.vel = v_out;
    ChainCall cc;
    cc.M = Mi; cc.T = T; cc.frames = frames_dev; cc.mod = mod; cc.fold = fold;
    cc.attn.B = B; cc.attn.R = c->R; cc.attn.P = c->P; cc.attn.ref_len = c->ref_len; cc.attn.ph_len = c->ph_len;
    cc.attn.kv_ref = c->kv_ref; cc.attn.kv_text = c->kv_text;
    cc.qkv_db = attn_mode != 0;
    CK(cudaMemsetAsync(ws.ready, 0, DenoiseWs::ready_ints(M) * sizeof(int), st));
    CK(chain_stats_cast(st, ws.x, Mi, mod + D /*scale_msa of block 0*/, ws.a, ws.stats));
    if (attn_mode == 1) {  // the whole evaluation in one launch
      cb.ready = ws.ready;
      cc.n_phases = 0;
      auto add = [&](int kind, int blk) { cc.kind[cc.n_phases] = kind; cc.blk[cc.n_phases] = blk; ++cc.n_phases; };
      add(CHAIN_QKVG, 0);
      for (int i = 0; i < NBLK; ++i) {
        add(CHAIN_ATTN, i); add(CHAIN_OUT, i); add(CHAIN_W13, i); add(CHAIN_W2, i);
        if (i + 1 < NBLK) add(CHAIN_QKVG, i + 1); else add(CHAIN_VEL, 0);
      }
      CK(launch_dit_chain(st, e->chain_w, cb, cc));
      return;
    }
    int launch = 0;
    auto run = [&](std::initializer_list<std::pair<int, int>> phases) {
      cb.ready = ws.ready + launch * ready_ints;
      if (e->chain_split) {  // one GEMM per launch: kernel boundaries instead of the in-kernel ready counters
        for (const auto& ph : phases) {
          cc.n_phases = 1; cc.kind[0] = ph.first; cc.blk[0] = ph.second;
          CK(cudaMemsetAsync(cb.ready, 0, ready_ints * sizeof(int), st));  // every launch claims its tiles from zero
          CK(launch_dit_chain(st, e->chain_w, cb, cc));
        }
      } else {
        cc.n_phases = 0;
        for (const auto& ph : phases) { cc.kind[cc.n_phases] = ph.first; cc.blk[cc.n_phases] = ph.second; ++cc.n_phases; }
        CK(launch_dit_chain(st, e->chain_w, cb, cc));
      }
      ++launch;
    };
    run({{CHAIN_QKVG, 0}});
    if (attn_mode == 2) {  // one launch per block
      for (int i = 0; i < NBLK; ++i) {
        run({{CHAIN_ATTN, i}, {CHAIN_OUT, i}, {CHAIN_W13, i}, {CHAIN_W2, i}, {i + 1 < NBLK ? CHAIN_QKVG : CHAIN_VEL, i + 1 < NBLK ? i + 1 : 0}});
      }
      return;
    }
    for (int i = 0; i < NBLK; ++i) {
      AttnSeg segs[3];
      segs[0].k = ws.kb; segs[0].v = ws.vb; segs[0].len = frames_dev; segs[0].n_max = T;
      segs[1].k = c->kv_ref + (2 * i) * c->ref_stride(); segs[1].v = c->kv_ref + (2 * i + 1) * c->ref_stride();
      segs[1].len = c->ref_len; segs[1].n_max = c->R;
      segs[2].k = c->kv_text + (2 * i) * c->text_stride(); segs[2].v = c->kv_text + (2 * i + 1) * c->text_stride();
      segs[2].len = c->ph_len; segs[2].n_max = c->P;
      static const bool skip_attn = [] { const char* v = getenv("STTS_DEBUG_SKIP_ATTENTION"); return v && v[0] == '1'; }();
      if (!skip_attn) {  // (timing experiments only: without it the output is garbage)
        CK(attention_bf16(st, ws.qb, B, T, H, HD, HDP, segs, 3, ws.qkvg, D, 0, ws.ob));
      }
      if (i + 1 < NBLK) run({{CHAIN_OUT, i}, {CHAIN_W13, i}, {CHAIN_W2, i}, {CHAIN_QKVG, i + 1}});
      else run({{CHAIN_OUT, i}, {CHAIN_W13, i}, {CHAIN_W2, i}, {CHAIN_VEL, 0}});
    }
    return;
  }
  for (int i = 0; i < NBLK; ++i) {
    const DitBlockW& w = e->blk[i];
    const float* m = mod + i * 6 * D;  // [shift_msa, scale_msa, tanh gate_msa, shift_mlp, scale_mlp, tanh gate_mlp]
    CK(ln_mod_bf16(st, ws.x, M, T, D, m + D, m, ld_mod, 1e-6f, ws.a));
    GemmEpi eq;
    eq.bias = w.bqkvg; eq.out_f32 = ws.qkvg; eq.ld_out = 4 * D;
    linear(e, ws.a, M, D, D, w.wqkvg, 4 * D, D, eq);
    CK(head_split_qkv_bf16(st, ws.qkvg, 4 * D, D, M, T, H, HD, HDP, w.qn, w.kn, 1e-6f, 64, e->cos64, e->sin64, ws.qb, ws.kb,
                           ws.vb));
    AttnSeg segs[3];
    segs[0].k = ws.kb; segs[0].v = ws.vb; segs[0].len = frames_dev; segs[0].n_max = T;
    segs[1].k = c->kv_ref + (2 * i) * c->ref_stride(); segs[1].v = c->kv_ref + (2 * i + 1) * c->ref_stride();
    segs[1].len = c->ref_len; segs[1].n_max = c->R;
    segs[2].k = c->kv_text + (2 * i) * c->text_stride(); segs[2].v = c->kv_text + (2 * i + 1) * c->text_stride();
    segs[2].len = c->ph_len; segs[2].n_max = c->P;
    CK(attention_bf16(st, ws.qb, B, T, H, HD, HDP, segs, 3, ws.qkvg, 4 * D, 3 * D, ws.ob));
    GemmEpi eo;  // x += tanh(gate_msa) * mask(to_out(o))   (dit.py:115-118,198)
    eo.row_len = frames_dev; eo.rows_per_batch = T; eo.rowgate = m + 2 * D; eo.ld_gate = ld_mod;
    eo.residual = ws.x; eo.ld_res = D; eo.out_f32 = ws.x; eo.ld_out = D;
    linear(e, ws.ob, M, H * HDP, H * HDP, w.wo, D, H * HDP, eo);
    CK(ln_mod_bf16(st, ws.x, M, T, D, m + 4 * D, m + 3 * D, ld_mod, 1e-6f, ws.a));
    GemmEpi e1;
    e1.bias = w.b13; e1.act = ACT_SWIGLU16; e1.out_bf16 = ws.hb; e1.ld_out = FF;
    linear(e, ws.a, M, D, D, w.w13, 2 * FF, D, e1);
    GemmEpi e2;  // x += tanh(gate_mlp) * ff   (dit.py:201)
    e2.bias = w.b2; e2.rows_per_batch = T; e2.rowgate = m + 5 * D; e2.ld_gate = ld_mod;
    e2.residual = ws.x; e2.ld_res = D; e2.out_f32 = ws.x; e2.ld_out = D;
    linear(e, ws.hb, M, FF, FF, w.w2, D, FF, e2);
  }
  const float* mf = mod + NBLK * 6 * D;  // [scale, shift] (dit.py:37)
  CK(ln_mod_bf16(st, ws.x, M, T, D, mf, mf + D, ld_mod, 1e-6f, ws.a));
  GemmEpi ev;
  ev.bias = e->vel_b; ev.out_f32 = v_out; ev.ld_out = LAT;
  linear(e, ws.a, M, D, D, e->vel_w, LAT, D, ev, 64);
}

void alpha_sigma(float t32, float* alpha, float* sigma) {  // infer/onnx.py:31-39 (numpy: fp64 inside, fp32 out)
  double t = static_cast<double>(t32);
  const double eps = 1e-5;
  t = t < eps ? eps : (t > 1 - eps ? 1 - eps : t);
  const double c = cos(M_PI / 2 * t);
  const double a2 = c * c;
  const double log_snr = log(a2 / (1 - a2)) + 2 * log(0.5);
  const double alpha_sq = 1.0 / (1.0 + exp(-log_snr));
  *alpha = static_cast<float>(sqrt(alpha_sq));
  *sigma = static_cast<float>(sqrt(1 - alpha_sq));
}

// DMD loop of infer/onnx.py:98-125 on device. noise_dev may be null (Philox). out: device [B,T,64].
std::vector<float> resolve_timesteps(int steps, const float* timesteps) {
  std::vector<float> ts(steps);
  for (int s = 0; s < steps; ++s) {  // default: np.linspace(1, 0, steps, dtype=float32) (infer/onnx.py:102)
    ts[s] = timesteps ? timesteps[s]
                      : (steps == 1 ? 1.0f : static_cast<float>(1.0 + (0.0 - 1.0) * static_cast<double>(s) / (steps - 1)));
  }
  return ts;
}

// Build (or fetch) the adaLN tables of every timestep.  Allocates and synchronises, so it runs BEFORE any capture.
std::vector<const float*> prepare_mods(stts_engine* e, const std::vector<float>& ts) {
  std::vector<const float*> mods(ts.size());
  for (size_t s = 0; s < ts.size(); ++s) mods[s] = cached_mod(e, ts[s]);
  return mods;
}

std::vector<const float*> prepare_folds(stts_engine* e, const std::vector<float>& ts) {
  std::vector<const float*> folds(ts.size());
  for (size_t s = 0; s < ts.size(); ++s) folds[s] = cached_fold(e, ts[s]);
  return folds;
}

// Launch-only (graph-capturable).  seed_dev: device u64 read by the Philox kernel when noise_dev is null.
void sample(stts_engine* e, const stts_cond* c, const int* frames_dev, int B, int T, const std::vector<float>& ts,
            const std::vector<const float*>& mods, const std::vector<const float*>& folds, const float* noise_dev,
            const unsigned long long* seed_dev, float* x_pred) {
  cudaStream_t st = e->st;
  const int steps = static_cast<int>(ts.size());
  const long long n = static_cast<long long>(B) * T * LAT;
  DenoiseWs ws;
  ws.alloc(st, static_cast<long long>(B) * T);
  Tmp<float> xt(st, n), v(st, n), nz;
  Tmp<bf16> xtb(st, n);
  if (noise_dev == nullptr) nz.alloc(st, n);
  CK(cudaMemsetAsync(x_pred, 0, n * sizeof(float), st));
  for (int s = 0; s < steps; ++s) {
    float alpha, sigma;
    alpha_sigma(ts[s], &alpha, &sigma);
    const float* nzs = noise_dev ? noise_dev + s * n : nz.p;
    if (noise_dev == nullptr) CK(philox_normal(st, seed_dev, static_cast<unsigned long long>(s), n, nz));
    CK(noise_mix(st, x_pred, nzs, alpha, sigma, n, xt, xtb));
    denoise(e, c, ws, xtb, frames_dev, mods[s], 0, folds[s], B, T, v);
    CK(dmd_update(st, xt, v, alpha, sigma, n, x_pred));
  }
}

// Teacher sampler (BASELINE config 5).  The reference has no inference script for its teacher; this is
// the sampler its distillation code implies: 3-way classifier-free guidance exactly as get_x_pred builds it
// (scripts/train/dmd2/distill.py:74-103: rows [cond | text dropped | speaker dropped], scales 2.0 / 1.5) and a
// deterministic DDIM walk over t = linspace(1, 0, steps + 1) in the v-parameterisation of train/utils.py:54-67:
//   x0 = a x - s v,  eps = s x + a v,  x' = a' x0 + s' eps = (a' a + s' s) x + (s' a - a' s) v.
// c3 holds the conditions of the 3B-row batch; frames3_dev its frame counts (frames repeated three times).
void sample_teacher(stts_engine* e, const stts_cond* c3, const int* frames3_dev, int B, int T, int steps, float cfg_text,
                    float cfg_spk, const float* noise_dev, const unsigned long long* seed_dev, float* x) {
  cudaStream_t st = e->st;
  const long long n = static_cast<long long>(B) * T * LAT;
  std::vector<float> ts(steps + 1);
  for (int s = 0; s <= steps; ++s) ts[s] = static_cast<float>(1.0 - static_cast<double>(s) / steps);
  const std::vector<const float*> mods = prepare_mods(e, std::vector<float>(ts.begin(), ts.end() - 1));
  const std::vector<const float*> folds = prepare_folds(e, std::vector<float>(ts.begin(), ts.end() - 1));
  DenoiseWs ws;
  ws.alloc(st, static_cast<long long>(3) * B * T);
  Tmp<float> v3(st, 3 * n);
  Tmp<bf16> xb(st, 3 * n);
  if (noise_dev != nullptr) {
    CK(cudaMemcpyAsync(x, noise_dev, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  } else {
    CK(philox_normal(st, seed_dev, 0ull, n, x));
  }
  for (int s = 0; s < steps; ++s) {
    float a, sg, an, sn;
    alpha_sigma(ts[s], &a, &sg);
    alpha_sigma(ts[s + 1], &an, &sn);
    CK(repeat_cast_bf16(st, x, n, 3, xb));
    denoise(e, c3, ws, xb, frames3_dev, mods[s], 0, folds[s], 3 * B, T, v3);
    CK(cfg_ddim_update(st, x, v3, n, cfg_text, cfg_spk, an * a + sn * sg, sn * a - an * sg));
  }
}

// ------------------------------------------------------------------ vocoder (hf:406-500)
struct VocWs {  // activations of one decode; persistent inside a Plan, temporaries otherwise
  float *xa = nullptr, *xb = nullptr;
  bf16 *a = nullptr, *hbuf = nullptr, *xh = nullptr, *latb = nullptr;
  static size_t act_elems(long long frames) { return static_cast<size_t>(frames) * 102400; }  // max rows*C per frame
};
enum VocPart : int { VOC_FRONT = 1, VOC_TAIL = 2, VOC_ALL = 3 };

// The ConvNeXt layers of one stage, in place on `cur` (hf:284-297); `oth` is scratch of the same size.  The last layer
// also writes a bf16 copy to ws.xh when the next operator is a GEMM over the activations.
void convnext_layers(stts_engine* e, const std::vector<VocLayerW>& layers, float*& cur, float*& oth, const VocWs& ws, int B,
                     int Ts, int C, bool bf16_copy) {
  cudaStream_t st = e->st;
  const long long M = static_cast<long long>(B) * Ts;
  const size_t nl = layers.size();
  if (C <= 64 && e->fused_tail) {
    // fused layers are out-of-place: ping-pong cur -> oth -> cur -> ...; `cur` is the result afterwards
    for (size_t l = 0; l < nl; ++l) {
      const VocLayerW& w = layers[l];
      bf16* hb = (bf16_copy && l + 1 == nl) ? ws.xh : nullptr;
      CK(convnext_fused(st, cur, B, Ts, C, w.norm_w, w.conv_w, w.conv_b, w.gamma, w.ffn_norm_w, w.w1, w.b1, w.w2h,
                        w.b2, w.ffn_gamma, 1e-5f, oth, hb));
      std::swap(cur, oth);
    }
    return;
  }
  for (size_t l = 0; l < nl; ++l) {
    const VocLayerW& w = layers[l];
    CK(convnext_mix(st, cur, B, Ts, C, w.norm_w, w.conv_w, w.conv_b, w.gamma, w.ffn_norm_w, 1e-5f, oth, ws.a));
    if (C == 128 && e->fused_ffn) {  // hidden activation stays on the SM
      CK(ffn_fused(st, ws.a, oth, M, C, w.w1, w.b1, w.w2h, w.b2, w.ffn_gamma, cur, (bf16_copy && l + 1 == nl) ? ws.xh : nullptr));
      continue;
    }
    GemmEpi e1;
    e1.bias = w.b1; e1.act = ACT_GELU; e1.gelu2_f16 = 1; e1.out_bf16 = ws.hbuf; e1.ld_out = 4 * C;
    linear(e, ws.a, M, C, C, w.w1, 4 * C, C, e1, 0, false, /*split_ok=*/true);
    GemmEpi e2;  // x = y + ffn_gamma * (W2 h + b2)   (hf:296-297)
    e2.bias = w.b2; e2.colscale = w.ffn_gamma; e2.residual = oth; e2.ld_res = C; e2.out_f32 = cur; e2.ld_out = C;
    if (bf16_copy && l + 1 == nl) e2.out_bf16 = ws.xh;  // bf16 copy feeds the next (transposed / strided) conv GEMM
    linear(e, ws.hbuf, M, 4 * C, 4 * C, static_cast<const bf16*>(w.w2h), C, 4 * C, e2, 0, /*f16=*/true, /*split_ok=*/true);
  }
}

// Launch-only (graph-capturable).  FRONT = stem + the ConvNeXt layers of the stages with C >= 256 and the upsamplers
// between them (tensor-bound); TAIL = the upsampler into C = 128, stages with C <= 128 and the head (HBM-bound: this is
// SURVEY 8(d)'s up3 + up4 + up5 + head group, whose 753.6 MB per 10 s utterance bench.py's roofline is quoted on).
// Stage s works in buffer `cur`; every transposed convolution writes `oth` and swaps.
void decode(stts_engine* e, const float* lat_dev, int B, int T, float* audio_dev, const VocWs& ws, int part) {
  cudaStream_t st = e->st;
  const long long frames = static_cast<long long>(B) * T;
  if (part & VOC_FRONT) {
    CK(cast_bf16(st, lat_dev, frames * LAT, ws.latb));
    // stem: causal Conv1d(64 -> 2048, k=7) as a 7-tap GEMM
    GemmShape s;
    s.B = B; s.T = T; s.N = 2048; s.K = 64; s.taps = 7; s.tap_shift0 = -6; s.tap_step = 1;
    GemmEpi ep;
    ep.bias = e->stem_b; ep.out_f32 = ws.xa; ep.ld_out = 2048;
    CK(launch_gemm(st, pick_bn(frames, 2048, 7), GemmA{ws.latb, LAT, LAT}, GemmW{e->stem_w, 2048, 7 * 64}, s, ep));
  }
  int Ts = T;
  float* cur = ws.xa;
  float* oth = ws.xb;
  for (int s = 0; s < 7; ++s) {
    const int C = VOC_C[s];
    const long long M = static_cast<long long>(B) * Ts;
    const bool layers_mine = (s < 4) ? (part & VOC_FRONT) != 0 : (part & VOC_TAIL) != 0;
    const bool up_mine = (s < 3) ? (part & VOC_FRONT) != 0 : (part & VOC_TAIL) != 0;
    if (layers_mine) {
      convnext_layers(e, e->voc[s], cur, oth, ws, B, Ts, C, /*bf16_copy=*/s < 6);
    } else if (C <= 64 && e->fused_tail && (VOC_DEPTH[s] & 1)) {
      std::swap(cur, oth);  // the other part's fused layers ping-pong between the buffers: follow them
    }
    if (s < 6) {  // causal ConvTranspose1d(k=2r, stride=r) as a 2-tap GEMM with N = r*Cout (hf:219-260)
      if (up_mine) {
        const int r = VOC_R[s], cout = VOC_C[s + 1];
        GemmShape g;
        g.B = B; g.T = Ts; g.N = r * cout; g.K = C; g.taps = 2; g.tap_shift0 = 0; g.tap_step = -1;
        GemmEpi ep;
        ep.bias = e->up_b[s]; ep.out_f32 = oth; ep.ld_out = r * cout;
        const int bn = pick_bn(M, r * cout, 2 * ((C + 63) / 64));
        set_splits(e, g, ep, bn);
        CK(launch_gemm(st, bn, GemmA{ws.xh, C, C}, GemmW{e->up_w[s], r * cout, 2 * C}, g, ep));
      }
      Ts *= VOC_R[s];
      std::swap(cur, oth);
    }
  }
  if (part & VOC_TAIL) CK(head_conv(st, cur, B, Ts, 32, e->head_w, e->head_b, audio_dev));
}

// Codec encoder (codec/onnx.py:56-75 == encoder.onnx; arithmetic per hf:300-403).  audio fp32 [B, N] with N a multiple
// of the hop (3200) -> latents fp32 [B, N/3200, 64] (the VAE mean).  Launch-only.
//   stem Conv1d(1->32,k7) | 3 ConvNeXt @32 | 6 x [strided causal Conv1d(C->2C, k=2r, stride r) + ConvNeXt layers] | head
// A strided conv with k = 2r is a two-tap GEMM over the activations regrouped r rows at a time ([T, C] read as
// [T/r, r*C]): out[t'] = W_a X'[t'-1] + W_b X'[t'] -- the mirror image of the decoder's transposed convolutions.
void encode_audio(stts_engine* e, const float* audio_dev, int B, int N, float* lat_dev, const VocWs& ws) {
  cudaStream_t st = e->st;
  float* cur = ws.xa;
  float* oth = ws.xb;
  CK(audio_stem_conv(st, audio_dev, B, N, 32, e->enc_stem_w, e->enc_stem_b, cur));
  int Ts = N;
  for (int s = 0; s < 7; ++s) {
    const int C = ENC_C[s];
    convnext_layers(e, e->enc[s], cur, oth, ws, B, Ts, C, /*bf16_copy=*/true);
    if (s < 6) {
      const int r = ENC_R[s], cout = ENC_C[s + 1];
      GemmShape g;
      g.B = B; g.T = Ts / r; g.N = cout; g.K = r * C; g.taps = 2; g.tap_shift0 = -1; g.tap_step = 1;
      GemmEpi ep;
      ep.bias = e->enc_down_b[s]; ep.out_f32 = oth; ep.ld_out = cout;
      CK(launch_gemm(st, pick_bn(static_cast<long long>(B) * (Ts / r), cout, 2 * ((r * C + 63) / 64)),
                     GemmA{ws.xh, r * C, r * C}, GemmW{e->enc_down_w[s], cout, 2 * r * C}, g, ep));
      std::swap(cur, oth);
      Ts /= r;
    }
  }
  GemmShape h;  // head: causal Conv1d(2048 -> 64, k=7) as a 7-tap GEMM
  h.B = B; h.T = Ts; h.N = LAT; h.K = 2048; h.taps = 7; h.tap_shift0 = -6; h.tap_step = 1;
  GemmEpi ep;
  ep.bias = e->enc_head_b; ep.out_f32 = lat_dev; ep.ld_out = LAT;
  CK(launch_gemm(st, 64, GemmA{ws.xh, 2048, 2048}, GemmW{e->enc_head_w, LAT, 7 * 2048}, h, ep));
}

struct VocTmp {  // stream-ordered temporaries for the stand-alone decode entry point
  Tmp<float> xa, xb;
  Tmp<bf16> a, hbuf, xh, latb;
  VocWs ws;
  VocTmp(cudaStream_t st, long long frames) {
    const size_t act = VocWs::act_elems(frames);
    xa.alloc(st, act); xb.alloc(st, act); a.alloc(st, act); hbuf.alloc(st, act * 4); xh.alloc(st, act);
    latb.alloc(st, frames * LAT);
    ws.xa = xa; ws.xb = xb; ws.a = a; ws.hbuf = hbuf; ws.xh = xh; ws.latb = latb;
  }
};

// ------------------------------------------------------------------ plans: persistent state + CUDA graphs per shape
// Everything stts_synthesize needs for one (B, R, P, T, steps) shape lives here, so the steady state performs no
// allocation and no intermediate synchronisation, and the ~850 kernel launches of a step replay as four graphs
// (condition encoder | DMD loop | vocoder front | vocoder tail) with CUDA events between them for stage timing.
struct Plan {
  int B = 0, R = 0, P = 0, T = 0, steps = 0;
  stts_cond cond;
  int* frames_dev = nullptr;
  int* h_lens = nullptr;  // pinned [3][B]: ref_len, ph_len, frames
  float *ref = nullptr, *noise = nullptr, *lat = nullptr, *audio = nullptr;
  long long* ids = nullptr;
  VocWs voc;
  std::vector<void*> owned;
  std::vector<float> ts;
  std::vector<const float*> mods, folds;
  cudaGraphExec_t g_cond = nullptr, g_sample[2] = {nullptr, nullptr}, g_front = nullptr, g_tail = nullptr;
  unsigned long long n_cond = 0, n_sample[2] = {0, 0}, n_front = 0, n_tail = 0;  // kernels per graph
  int graph_state = 0;  // 0 = not captured yet, 1 = captured, -1 = capture failed: stay eager
  uint64_t last_use = 0;
  template <typename T>
  T* dalloc(size_t n) {
    T* p = nullptr;
    CK(cudaMalloc(reinterpret_cast<void**>(&p), (n ? n : 1) * sizeof(T)));
    owned.push_back(p);
    return p;
  }
  void destroy() {
    for (cudaGraphExec_t g : {g_cond, g_sample[0], g_sample[1], g_front, g_tail}) if (g) cudaGraphExecDestroy(g);
    for (void* p : owned) cudaFree(p);
    if (h_lens) cudaFreeHost(h_lens);
    owned.clear();
  }
};

void set_seed(stts_engine* e, uint64_t seed) {
  *e->seed_host = seed;
  CK(cudaMemcpyAsync(e->seed_dev, e->seed_host, sizeof(unsigned long long), cudaMemcpyHostToDevice, e->st));
}

Plan* get_plan(stts_engine* e, int B, int R, int P, int T, int steps) {
  const std::array<int, 5> key = {B, R, P, T, steps};
  auto it = e->plans.find(key);
  if (it != e->plans.end()) {
    it->second->last_use = ++e->use_counter;
    return it->second;
  }
  if (e->plans.size() >= 6) {  // bound the memory held by cached shapes: evict the least recently used plan
    auto victim = e->plans.begin();
    for (auto j = e->plans.begin(); j != e->plans.end(); ++j) {
      if (j->second->last_use < victim->second->last_use) victim = j;
    }
    CK(cudaStreamSynchronize(e->st));
    victim->second->destroy();
    delete victim->second;
    e->plans.erase(victim);
  }
  Plan* p = new Plan();
  try {
    p->B = B; p->R = R; p->P = P; p->T = T; p->steps = steps;
    const long long n = static_cast<long long>(B) * T * LAT, frames = static_cast<long long>(B) * T;
    p->cond.B = B; p->cond.R = R; p->cond.P = P;
    p->cond.ref_len = p->dalloc<int>(B);
    p->cond.ph_len = p->dalloc<int>(B);
    p->frames_dev = p->dalloc<int>(B);
    p->cond.kv_ref = p->dalloc<bf16>(NBLK * 2 * p->cond.ref_stride());
    p->cond.kv_text = p->dalloc<bf16>(NBLK * 2 * p->cond.text_stride());
    p->ref = p->dalloc<float>(static_cast<size_t>(B) * R * LAT);
    p->ids = p->dalloc<long long>(static_cast<size_t>(B) * P);
    p->noise = p->dalloc<float>(static_cast<size_t>(steps) * n);
    p->lat = p->dalloc<float>(n);
    p->audio = p->dalloc<float>(static_cast<size_t>(frames) * STTS_HOP_SIZE);
    const size_t act = VocWs::act_elems(frames);
    p->voc.xa = p->dalloc<float>(act); p->voc.xb = p->dalloc<float>(act);
    p->voc.a = p->dalloc<bf16>(act); p->voc.hbuf = p->dalloc<bf16>(act * 4); p->voc.xh = p->dalloc<bf16>(act);
    p->voc.latb = p->dalloc<bf16>(frames * LAT);
    CK(cudaHostAlloc(reinterpret_cast<void**>(&p->h_lens), 3 * B * sizeof(int), cudaHostAllocDefault));
    p->ts = resolve_timesteps(steps, nullptr);
    p->mods = prepare_mods(e, p->ts);
    p->folds = prepare_folds(e, p->ts);
  } catch (...) {
    p->destroy();
    delete p;
    throw;
  }
  p->last_use = ++e->use_counter;
  e->plans[key] = p;
  return p;
}

// Record `fn`'s launches into an executable graph (nothing runs).  Relaxed mode: the launchers make driver calls
// (tensor-map encoding) that are not stream operations.
template <typename F>
cudaGraphExec_t capture(stts_engine* e, F&& fn, unsigned long long* n_kernels) {
  cudaGraph_t graph = nullptr;
  CK(cudaStreamBeginCapture(e->st, cudaStreamCaptureModeRelaxed));
  ++tl_capturing;  // launches below are recorded, not executed: the launch counter ignores them
  try {
    fn();
  } catch (...) {
    --tl_capturing;
    cudaStreamEndCapture(e->st, &graph);
    if (graph) cudaGraphDestroy(graph);
    throw;
  }
  --tl_capturing;
  CK(cudaStreamEndCapture(e->st, &graph));
  {  // kernels per replay = kernel nodes of the graph
    size_t n = 0;
    CK(cudaGraphGetNodes(graph, nullptr, &n));
    std::vector<cudaGraphNode_t> nodes(n);
    if (n) CK(cudaGraphGetNodes(graph, nodes.data(), &n));
    unsigned long long k = 0;
    for (size_t i = 0; i < n; ++i) {
      cudaGraphNodeType ty;
      CK(cudaGraphNodeGetType(nodes[i], &ty));
      if (ty == cudaGraphNodeTypeKernel) ++k;
    }
    *n_kernels = k;
  }
  cudaGraphExec_t exec = nullptr;
  cudaError_t err = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (err != cudaSuccess) throw Err(STTS_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(err));
  return exec;
}

}  // namespace

// ==================================================================== C ABI

namespace {
thread_local std::string g_create_err;

int guard_impl(stts_engine* e, const std::function<void()>& fn) {
  try {
    if (e) CK(cudaSetDevice(e->device));
    fn();
    return STTS_OK;
  } catch (const Err& ex) {
    (e ? e->err : g_create_err) = ex.what();
    cudaGetLastError();
    return ex.code;
  } catch (const std::exception& ex) {
    (e ? e->err : g_create_err) = ex.what();
    return STTS_ERR_INVALID;
  }
}
// ------------------------------------------------------------------ resampler (infer/utils.py:7-23)
double bessel_i0(double x) {  // sum_k ((x/2)^k / k!)^2
  const double q = 0.25 * x * x;
  double term = 1.0, sum = 1.0;
  for (int k = 1; k < 500; ++k) {
    term *= q / (static_cast<double>(k) * k);
    sum += term;
    if (term < 1e-17 * sum) break;
  }
  return sum;
}

int gcd_int(int a, int b) {
  while (b) {
    const int t = a % b;
    a = b;
    b = t;
  }
  return a;
}

// torchaudio.functional._get_sinc_resample_kernel with the reference's arguments (resampling_method
// "sinc_interp_kaiser", lowpass_filter_width 1024, rolloff 0.94, beta 14.769656459379492), evaluated in fp64 and
// rounded to fp32 like torchaudio does (dtype=None).  Two fp32 roundings of the original are kept: the phase offsets
// -j/up come from an integer tensor divided in fp32, and beta is a fp32 tensor.
const stts_engine::ResampleBank& resample_bank(stts_engine* e, int sr_from, int sr_to) {
  const int g = gcd_int(sr_from, sr_to);
  const int down = sr_from / g, up = sr_to / g;
  auto it = e->banks.find({down, up});
  if (it != e->banks.end()) return it->second;
  const double lpw = 1024.0, rolloff = 0.94;
  const double beta = static_cast<double>(static_cast<float>(14.769656459379492));
  const double base_freq = static_cast<double>(down < up ? down : up) * rolloff;
  const int width = static_cast<int>(ceil(lpw * down / base_freq));
  const long long K = 2LL * width + down;
  const long long smem = (31LL * down + K) * 4;
  if (smem > 200 * 1024 || K * up > (64LL << 20)) {
    throw Err(STTS_ERR_INVALID, "unsupported sample-rate pair: the reduced ratio " + std::to_string(down) + ":" +
                                    std::to_string(up) + " needs a " + std::to_string(K) + "-tap x " + std::to_string(up) +
                                    "-phase filter bank");
  }
  std::vector<float> h(static_cast<size_t>(K) * up);
  const double i0b = bessel_i0(beta), scale = base_freq / down, pi = 3.14159265358979323846;
  for (int j = 0; j < up; ++j) {
    const double phase = static_cast<double>(static_cast<float>(-j) / static_cast<float>(up));
    for (long long k = 0; k < K; ++k) {
      double t = (phase + static_cast<double>(k - width) / down) * base_freq;
      t = t < -lpw ? -lpw : (t > lpw ? lpw : t);
      const double r = t / lpw;
      const double win = bessel_i0(beta * sqrt(1.0 - r * r)) / i0b;
      t *= pi;
      const double sinc = t == 0.0 ? 1.0 : sin(t) / t;
      h[static_cast<size_t>(j) * K + k] = static_cast<float>(sinc * win * scale);
    }
  }
  stts_engine::ResampleBank bk;
  bk.down = down;
  bk.up = up;
  bk.width = width;
  bk.K = static_cast<int>(K);
  bk.d = e->dalloc<float>(h.size(), false);
  CK(cudaMemcpyAsync(bk.d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, e->st));
  CK(cudaStreamSynchronize(e->st));  // h goes out of scope
  return e->banks.emplace(std::make_pair(down, up), bk).first->second;
}

void need_ready(stts_engine* e, bool dit = false, bool decoder = false) {
  if (!e->finalized) throw Err(STTS_ERR_WEIGHTS, "weights not finalized: call stts_finalize_weights first");
  if (dit && !e->has_dit) throw Err(STTS_ERR_WEIGHTS, "DiT weights (model 0) were not loaded");
  if (decoder && !e->has_decoder) throw Err(STTS_ERR_WEIGHTS, "codec decoder weights (model 1) were not loaded");
}
}  // namespace

namespace {
// Streams, events and the seed staging buffers: what every handle (created or cloned) owns for itself.
void init_instance(stts_engine* e) {
  CK(cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&e->st2, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
  for (auto& ev : e->ev) CK(cudaEventCreate(&ev));
  CK(cudaEventCreate(&e->ev_stop));
  CK(cudaMalloc(reinterpret_cast<void**>(&e->seed_dev), sizeof(unsigned long long)));
  CK(cudaHostAlloc(reinterpret_cast<void**>(&e->seed_host), sizeof(unsigned long long), cudaHostAllocDefault));
}
}  // namespace

extern "C" {

int stts_create(const stts_config* cfg, stts_engine** out) {
  if (!out) return STTS_ERR_INVALID;
  *out = nullptr;
  stts_engine* e = new stts_engine();
  int rc = guard_impl(nullptr, [&] {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
      throw Err(STTS_ERR_NO_DEVICE, "no CUDA device: smalltts_b200 has no CPU fallback");
    }
    e->device = cfg ? cfg->device : 0;
    if (e->device < 0 || e->device >= n) throw Err(STTS_ERR_INVALID, "bad device ordinal");
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, e->device));
    if (p.major != 10) {
      throw Err(STTS_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(p.major * 10 + p.minor) +
                                        ", this engine is built for sm_100a (B200) only");
    }
    CK(cudaSetDevice(e->device));
    init_instance(e);
    const char* ng = getenv("STTS_NO_GRAPH");
    e->use_graphs = !(ng && ng[0] == '1');
    const char* nf = getenv("STTS_NO_FUSED_TAIL");
    e->fused_tail = !(nf && nf[0] == '1');
    const char* nff = getenv("STTS_NO_FUSED_FFN");
    e->fused_ffn = !(nff && nff[0] == '1');
    const char* ns = getenv("STTS_SPLITK");
    e->use_split = ns && ns[0] == '1';
    const char* nc = getenv("STTS_NO_CHAIN");
    e->use_chain = !(nc && nc[0] == '1');
    const char* cs = getenv("STTS_CHAIN_SPLIT");
    e->chain_split = cs && cs[0] == '1';
    const char* ca = getenv("STTS_CHAIN_ATTN");
    if (ca && ca[0] >= '0' && ca[0] <= '2') e->chain_attn = ca[0] - '0';
    cudaMemPool_t pool;
    CK(cudaDeviceGetDefaultMemPool(&pool, e->device));
    uint64_t thr = UINT64_MAX;
    CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  });
  if (rc != STTS_OK) {
    delete e;
    return rc;
  }
  *out = e;
  return STTS_OK;
}

void stts_destroy(stts_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->st);
  if (e->owns_weights) {
    for (auto& m : e->raw) for (auto& kv : m) cudaFree(kv.second.d);
  }
  for (void* p : e->owned) cudaFree(p);
  for (auto& kv : e->plans) {
    kv.second->destroy();
    delete kv.second;
  }
  cudaFree(e->seed_dev);
  cudaFreeHost(e->seed_host);
  for (auto& ev : e->ev) cudaEventDestroy(ev);
  cudaEventDestroy(e->ev_stop);
  cudaEventDestroy(e->ev_fork);
  cudaEventDestroy(e->ev_join);
  cudaStreamDestroy(e->st2);
  cudaStreamDestroy(e->st);
  delete e;
}

int stts_engine_clone(stts_engine* src, stts_engine** out) {
  if (!src || !out) return STTS_ERR_INVALID;
  *out = nullptr;
  stts_engine* c = nullptr;
  const int rc = guard_impl(src, [&] {
    need_ready(src);
    CK(cudaStreamSynchronize(src->st));
    c = new stts_engine(*src);  // every packed-weight pointer and the raw-tensor table, by value
    c->owns_weights = false;
    c->owned.clear();      // what the clone allocates from here on (adaLN tables, filter banks) is its own
    c->plans.clear();
    c->mod_cache.clear();
    c->fold_cache.clear();
    c->split_scratch = nullptr;  // scratch is per handle (handles run concurrently): the clone gets its own below
    c->split_cnt = nullptr;
    c->banks.clear();
    c->err.clear();
    c->use_counter = 0;
    c->test_async = false;
    c->timing = stts_timing{0, 0, 0, 0, 0};
    c->voc_ms[0] = c->voc_ms[1] = 0;
    c->st = c->st2 = nullptr;
    c->ev_fork = c->ev_join = c->ev_stop = nullptr;
    for (auto& ev : c->ev) ev = nullptr;
    c->seed_dev = nullptr;
    c->seed_host = nullptr;
    init_instance(c);
    if (c->has_decoder) {
      c->split_scratch = c->dalloc<float>(static_cast<size_t>(c->split_scratch_floats), false);
      c->split_cnt = c->dalloc<int>(static_cast<size_t>(c->split_cnt_ints), true);
    }
  });
  if (rc != STTS_OK) {
    if (c) stts_destroy(c);  // tolerates the handles init_instance did not get to
    return rc;
  }
  *out = c;
  return STTS_OK;
}

const char* stts_last_error(const stts_engine* e) { return e ? e->err.c_str() : g_create_err.c_str(); }

int stts_load_weight(stts_engine* e, int model, const char* name, const float* data, int ndim, const int64_t* shape) {
  if (!e) return STTS_ERR_INVALID;
  return guard_impl(e, [&] {
    if (model < 0 || model > 2 || !name || !data || ndim < 0 || ndim > 4) throw Err(STTS_ERR_INVALID, "bad weight args");
    if (!e->owns_weights) throw Err(STTS_ERR_WEIGHTS, "this handle is a clone: load weights through its source engine");
    RawTensor t;
    t.numel = 1;
    for (int i = 0; i < ndim; ++i) {
      t.shape.push_back(shape[i]);
      t.numel *= static_cast<size_t>(shape[i]);
    }
    // 16-byte aligned rows are assumed by the float4 paths; cudaMalloc gives 256-byte alignment
    CK(cudaMalloc(reinterpret_cast<void**>(&t.d), (t.numel ? t.numel : 1) * sizeof(float)));
    CK(cudaMemcpyAsync(t.d, data, t.numel * sizeof(float), cudaMemcpyHostToDevice, e->st));
    CK(cudaStreamSynchronize(e->st));
    auto it = e->raw[model].find(name);
    if (it != e->raw[model].end()) cudaFree(it->second.d);
    e->raw[model][name] = t;
    e->finalized = false;
  });
}

int stts_finalize_weights(stts_engine* e) {
  if (!e) return STTS_ERR_INVALID;
  return guard_impl(e, [&] {
    if (e->finalized) return;
    if (e->packed_once) throw Err(STTS_ERR_WEIGHTS, "weights were already finalized once; create a new engine");
    e->packed_once = true;
    finalize(e);
  });
}

int stts_encode_conditions(stts_engine* e, const float* ref, const int64_t* ref_len, const int64_t* phonemes,
                           const int64_t* ph_len, int B, int R, int P, int mem, stts_cond** out) {
  if (!e || !out) return STTS_ERR_INVALID;
  *out = nullptr;
  return guard_impl(e, [&] {
    need_ready(e, true);
    *out = encode_conditions(e, ref, ref_len, phonemes, ph_len, B, R, P, mem);
  });
}

void stts_cond_free(stts_engine* e, stts_cond* c) {
  if (!c) return;
  if (e) {
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->st);
  }
  cudaFree(c->ref_len); cudaFree(c->ph_len); cudaFree(c->kv_ref); cudaFree(c->kv_text);
  delete c;
}

int stts_cond_read_kv(stts_engine* e, const stts_cond* c, int layer, int which, float* dst) {
  if (!e || !c || !dst || layer < 0 || layer >= NBLK || which < 0 || which > 3) return STTS_ERR_INVALID;
  return guard_impl(e, [&] {
    const bool is_ref = which < 2;
    const int N = is_ref ? c->R : c->P;
    const size_t stride = is_ref ? c->ref_stride() : c->text_stride();
    const bf16* src = (is_ref ? c->kv_ref : c->kv_text) + (2 * layer + (which & 1)) * stride;
    std::vector<bf16> h(stride);
    CK(cudaMemcpy(h.data(), src, stride * sizeof(bf16), cudaMemcpyDeviceToHost));
    for (int b = 0; b < c->B; ++b)
      for (int hh = 0; hh < H; ++hh)
        for (int n = 0; n < N; ++n)
          for (int d = 0; d < HD; ++d)
            dst[((static_cast<size_t>(b) * H + hh) * N + n) * HD + d] =
                op16_to_float(h[((static_cast<size_t>(b) * N + n) * H + hh) * HDP + d]);
  });
}

int stts_denoise_step(stts_engine* e, const stts_cond* c, const float* x_t, const int64_t* frames, const float* t,
                      int B, int T, int mem, float* velocity) {
  if (!e || !c || !x_t || !frames || !t || !velocity) return STTS_ERR_INVALID;
  return guard_impl(e, [&] {
    need_ready(e, true);
    if (B != c->B || T < 1 || T > ROPE_MAX) throw Err(STTS_ERR_INVALID, "batch/T mismatch with conditions");
    cudaStream_t st = e->st;
    const long long n = static_cast<long long>(B) * T * LAT;
    Tmp<int> fr;
    lens_to_dev(e, frames, B, T, fr);
    Tmp<float> hx, vout;
    const float* xd = to_dev(e, x_t, n, mem, hx);
    Tmp<bf16> xb(st, n);
    CK(cast_bf16(st, xd, n, xb));
    bool same = true;
    for (int b = 1; b < B; ++b) same = same && (t[b] == t[0]);
    Tmp<float> table, td;
    const float* mod;
    int ld_mod = 0;
    const float* fold = nullptr;
    if (same) {
      mod = cached_mod(e, t[0]);
      fold = cached_fold(e, t[0]);
    } else {
      td.alloc(st, B);
      table.alloc(st, static_cast<size_t>(B) * MOD_LD);
      CK(cudaMemcpyAsync(td.p, t, B * sizeof(float), cudaMemcpyHostToDevice, st));
      compute_mod(e, td, B, table);
      mod = table;
      ld_mod = MOD_LD;
    }
    DenoiseWs ws;
    ws.alloc(st, static_cast<long long>(B) * T);
    float* vd = velocity;
    if (mem == STTS_MEM_HOST) {
      vout.alloc(st, n);
      vd = vout;
    }
    denoise(e, c, ws, xb, fr, mod, ld_mod, fold, B, T, vd);
    if (mem == STTS_MEM_HOST) CK(cudaMemcpyAsync(velocity, vd, n * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  });
}

int stts_sample(stts_engine* e, const stts_cond* c, const int64_t* frames, int B, int T, int steps,
                const float* timesteps, const float* noise, uint64_t seed, int mem, float* out_latents) {
  if (!e || !c || !frames || !out_latents) return STTS_ERR_INVALID;
  return guard_impl(e, [&] {
    need_ready(e, true);
    if (B != c->B || T < 1 || T > ROPE_MAX || steps < 1) throw Err(STTS_ERR_INVALID, "bad sample arguments");
    cudaStream_t st = e->st;
    const long long n = static_cast<long long>(B) * T * LAT;
    Tmp<int> fr;
    lens_to_dev(e, frames, B, T, fr);
    const std::vector<float> ts = resolve_timesteps(steps, timesteps);
    const std::vector<const float*> mods = prepare_mods(e, ts);
    const std::vector<const float*> folds = prepare_folds(e, ts);
    set_seed(e, seed);
    Tmp<float> hn, xo;
    const float* nd = noise ? to_dev(e, noise, static_cast<size_t>(steps) * n, mem, hn) : nullptr;
    float* xd = out_latents;
    if (mem == STTS_MEM_HOST) {
      xo.alloc(st, n);
      xd = xo;
    }
    sample(e, c, fr, B, T, ts, mods, folds, nd, e->seed_dev, xd);
    if (mem == STTS_MEM_HOST) CK(cudaMemcpyAsync(out_latents, xd, n * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  });
}

int stts_sample_teacher(stts_engine* e, const stts_cond* cond3, const int64_t* frames, int B, int T, int steps,
                        float cfg_scale_text, float cfg_scale_speaker, const float* noise, uint64_t seed, int mem,
                        float* out_latents) {
  if (!e || !cond3 || !frames || !out_latents) return STTS_ERR_INVALID;
  return guard_impl(e, [&] {
    need_ready(e, true);
    if (B < 1 || cond3->B != 3 * B || T < 1 || T > ROPE_MAX || steps < 1 || steps > 4096) {
      throw Err(STTS_ERR_INVALID, "bad teacher-sampler arguments (conditions must hold 3*B rows: cond | no text | no speaker)");
    }
    cudaStream_t st = e->st;
    const long long n = static_cast<long long>(B) * T * LAT;
    std::vector<int64_t> f3(3 * B);
    for (int i = 0; i < 3 * B; ++i) f3[i] = frames[i % B];
    Tmp<int> fr;
    lens_to_dev(e, f3.data(), 3 * B, T, fr);
    set_seed(e, seed);
    Tmp<float> hn, xo;
    const float* nd = noise ? to_dev(e, noise, static_cast<size_t>(n), mem, hn) : nullptr;
    float* xd = out_latents;
    if (mem == STTS_MEM_HOST) {
      xo.alloc(st, n);
      xd = xo;
    }
    sample_teacher(e, cond3, fr, B, T, steps, cfg_scale_text, cfg_scale_speaker, nd, e->seed_dev, xd);
    if (mem == STTS_MEM_HOST) CK(cudaMemcpyAsync(out_latents, xd, n * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  });
}

int stts_decode(stts_engine* e, const float* latents, int B, int T, int mem, float* audio) {
  if (!e || !latents || !audio) return STTS_ERR_INVALID;
  return guard_impl(e, [&] {
    need_ready(e, false, true);
    if (B < 1 || T < 1) throw Err(STTS_ERR_INVALID, "bad decode arguments");
    cudaStream_t st = e->st;
    const long long n = static_cast<long long>(B) * T * LAT;
    const long long na = static_cast<long long>(B) * T * STTS_HOP_SIZE;
    Tmp<float> hl, ao;
    const float* ld = to_dev(e, latents, n, mem, hl);
    float* ad = audio;
    if (mem == STTS_MEM_HOST) {
      ao.alloc(st, na);
      ad = ao;
    }
    VocTmp vt(st, static_cast<long long>(B) * T);
    CK(cudaEventRecord(e->ev[4], st));
    decode(e, ld, B, T, ad, vt.ws, VOC_FRONT);
    CK(cudaEventRecord(e->ev[5], st));
    decode(e, ld, B, T, ad, vt.ws, VOC_TAIL);
    CK(cudaEventRecord(e->ev[6], st));
    if (mem == STTS_MEM_HOST) CK(cudaMemcpyAsync(audio, ad, na * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&e->voc_ms[1], e->ev[4], e->ev[5]));
    CK(cudaEventElapsedTime(&e->voc_ms[0], e->ev[5], e->ev[6]));
  });
}

int stts_encode_audio(stts_engine* e, const float* audio, int B, int N, int mem, float* latents) {
  if (!e || !audio || !latents) return STTS_ERR_INVALID;
  return guard_impl(e, [&] {
    need_ready(e);
    if (!e->has_encoder) throw Err(STTS_ERR_WEIGHTS, "codec encoder weights (model 2) were not loaded");
    if (B < 1 || N < STTS_HOP_SIZE || (N % STTS_HOP_SIZE) != 0) {
      throw Err(STTS_ERR_INVALID, "audio length must be a positive multiple of the hop size (3200 samples)");
    }
    cudaStream_t st = e->st;
    const int T = N / STTS_HOP_SIZE;
    const long long na = static_cast<long long>(B) * N, nl = static_cast<long long>(B) * T * LAT;
    Tmp<float> ha, lo;
    const float* ad = to_dev(e, audio, static_cast<size_t>(na), mem, ha);
    float* ld = latents;
    if (mem == STTS_MEM_HOST) {
      lo.alloc(st, nl);
      ld = lo;
    }
    VocTmp vt(st, static_cast<long long>(B) * T);
    CK(cudaEventRecord(e->ev[4], st));
    encode_audio(e, ad, B, N, ld, vt.ws);
    CK(cudaEventRecord(e->ev[5], st));
    if (mem == STTS_MEM_HOST) CK(cudaMemcpyAsync(latents, ld, nl * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&e->timing.codec_enc_ms, e->ev[4], e->ev[5]));
  });
}

int64_t stts_resample_length(int N, int sr_from, int sr_to) {
  if (N < 0 || sr_from < 1 || sr_to < 1) return -1;
  const int g = gcd_int(sr_from, sr_to);
  const long long down = sr_from / g, up = sr_to / g;
  return (up * static_cast<long long>(N) + down - 1) / down;  // ceil(up * N / down), torchaudio target_length
}

int stts_resample(stts_engine* e, const float* audio, int B, int N, int sr_from, int sr_to, int mem, float* out) {
  if (!e || !audio || !out) return STTS_ERR_INVALID;
  return guard_impl(e, [&] {
    if (B < 1 || N < 1 || sr_from < 1 || sr_to < 1) throw Err(STTS_ERR_INVALID, "bad resample arguments");
    cudaStream_t st = e->st;
    const long long n_out = stts_resample_length(N, sr_from, sr_to);
    if (n_out < 1 || n_out > 0x7fffffffLL) throw Err(STTS_ERR_INVALID, "resampled length out of range");
    const size_t n_in_all = static_cast<size_t>(B) * N, n_out_all = static_cast<size_t>(B) * n_out;
    if (sr_from == sr_to) {  // resample_hq returns its input (infer/utils.py:20-21)
      CK(cudaMemcpyAsync(out, audio, n_in_all * sizeof(float), cudaMemcpyDefault, st));
      CK(cudaStreamSynchronize(st));
      return;
    }
    const stts_engine::ResampleBank& bk = resample_bank(e, sr_from, sr_to);
    Tmp<float> hin, hout;
    const float* xd = to_dev(e, audio, n_in_all, mem, hin);
    float* od = out;
    if (mem == STTS_MEM_HOST) {
      hout.alloc(st, n_out_all);
      od = hout;
    }
    CK(resample(st, xd, B, N, bk.d, bk.K, bk.width, bk.down, bk.up, od, static_cast<int>(n_out)));
    if (mem == STTS_MEM_HOST) CK(cudaMemcpyAsync(out, od, n_out_all * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  });
}

int stts_synthesize(stts_engine* e, const float* ref, const int64_t* ref_len, const int64_t* phonemes,
                    const int64_t* ph_len, const int64_t* frames, int B, int R, int P, int T, int steps,
                    const float* timesteps, const float* noise, uint64_t seed, int mem, float* audio) {
  if (!e || !ref || !ref_len || !phonemes || !ph_len || !frames || !audio) return STTS_ERR_INVALID;
  return guard_impl(e, [&] {
    need_ready(e, true, true);
    if (B < 1 || R < 1 || P < 1 || R > ROPE_MAX || P > ROPE_MAX || T < 1 || T > ROPE_MAX || steps < 1) {
      throw Err(STTS_ERR_INVALID, "bad synthesize arguments");
    }
    cudaStream_t st = e->st;
    const long long n = static_cast<long long>(B) * T * LAT;
    const long long na = static_cast<long long>(B) * T * STTS_HOP_SIZE;
    Plan* p = get_plan(e, B, R, P, T, steps);
    // ---- stage inputs into the plan's persistent buffers (async; no intermediate synchronisation)
    for (int b = 0; b < B; ++b) {
      if (ref_len[b] < 0 || ref_len[b] > R || ph_len[b] < 0 || ph_len[b] > P || frames[b] < 1 || frames[b] > T) {
        throw Err(STTS_ERR_INVALID, "length out of range");
      }
      p->h_lens[b] = static_cast<int>(ref_len[b]);
      p->h_lens[B + b] = static_cast<int>(ph_len[b]);
      p->h_lens[2 * B + b] = static_cast<int>(frames[b]);
    }
    CK(cudaMemcpyAsync(p->cond.ref_len, p->h_lens, B * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(p->cond.ph_len, p->h_lens + B, B * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(p->frames_dev, p->h_lens + 2 * B, B * sizeof(int), cudaMemcpyHostToDevice, st));
    const cudaMemcpyKind in_kind = mem == STTS_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    CK(cudaMemcpyAsync(p->ref, ref, static_cast<size_t>(B) * R * LAT * sizeof(float), in_kind, st));
    CK(cudaMemcpyAsync(p->ids, phonemes, static_cast<size_t>(B) * P * sizeof(long long), in_kind, st));
    if (noise) CK(cudaMemcpyAsync(p->noise, noise, static_cast<size_t>(steps) * n * sizeof(float), in_kind, st));
    set_seed(e, seed);
    // custom timestep schedules run eagerly; the default schedule is what the graphs were captured with
    std::vector<float> ts = resolve_timesteps(steps, timesteps);
    const bool default_ts = (ts == p->ts);
    std::vector<const float*> mods = default_ts ? p->mods : prepare_mods(e, ts);
    std::vector<const float*> folds = default_ts ? p->folds : prepare_folds(e, ts);
    const float* nd = noise ? p->noise : nullptr;

    auto run_cond = [&] { encode_conditions_core(e, &p->cond, p->ref, p->ids); };
    auto run_sample = [&](const float* nz) { sample(e, &p->cond, p->frames_dev, B, T, ts, mods, folds, nz, e->seed_dev, p->lat); };
    auto run_front = [&] { decode(e, p->lat, B, T, p->audio, p->voc, VOC_FRONT); };
    auto run_tail = [&] { decode(e, p->lat, B, T, p->audio, p->voc, VOC_TAIL); };

    const bool graphs = e->use_graphs && default_ts && p->graph_state == 1;
    CK(cudaEventRecord(e->ev[0], st));
    if (graphs) g_launch_count.fetch_add(p->n_cond + p->n_sample[noise ? 1 : 0] + p->n_front + p->n_tail, std::memory_order_relaxed);
    if (graphs) CK(cudaGraphLaunch(p->g_cond, st)); else run_cond();
    CK(cudaEventRecord(e->ev[1], st));
    if (graphs) CK(cudaGraphLaunch(p->g_sample[noise ? 1 : 0], st)); else run_sample(nd);
    CK(cudaEventRecord(e->ev[2], st));
    if (graphs) CK(cudaGraphLaunch(p->g_front, st)); else run_front();
    CK(cudaEventRecord(e->ev[5], st));
    if (graphs) CK(cudaGraphLaunch(p->g_tail, st)); else run_tail();
    CK(cudaEventRecord(e->ev[6], st));
    // the graphs write the plan's own audio buffer; hand the result to the caller
    const cudaMemcpyKind out_kind = mem == STTS_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    CK(cudaMemcpyAsync(audio, p->audio, na * sizeof(float), out_kind, st));
    CK(cudaEventRecord(e->ev[3], st));
    CK(cudaStreamSynchronize(st));
    e->timing.codec_enc_ms = 0.f;
    CK(cudaEventElapsedTime(&e->timing.cond_enc_ms, e->ev[0], e->ev[1]));
    CK(cudaEventElapsedTime(&e->timing.denoise_ms, e->ev[1], e->ev[2]));
    CK(cudaEventElapsedTime(&e->timing.codec_dec_ms, e->ev[2], e->ev[3]));
    CK(cudaEventElapsedTime(&e->timing.total_ms, e->ev[0], e->ev[3]));
    CK(cudaEventElapsedTime(&e->voc_ms[1], e->ev[2], e->ev[5]));
    CK(cudaEventElapsedTime(&e->voc_ms[0], e->ev[5], e->ev[6]));

    // First (eager) run of this shape succeeded: capture the four stage graphs for every later call.
    if (e->use_graphs && p->graph_state == 0 && default_ts) {
      try {
        p->g_cond = capture(e, run_cond, &p->n_cond);
        p->g_sample[0] = capture(e, [&] { run_sample(nullptr); }, &p->n_sample[0]);
        p->g_sample[1] = capture(e, [&] { run_sample(p->noise); }, &p->n_sample[1]);
        p->g_front = capture(e, run_front, &p->n_front);
        p->g_tail = capture(e, run_tail, &p->n_tail);
        p->graph_state = 1;
      } catch (const Err&) {
        p->graph_state = -1;  // keep working eagerly
        cudaGetLastError();
      }
    }
  });
}

int stts_get_timings(const stts_engine* e, stts_timing* out) {
  if (!e || !out) return STTS_ERR_INVALID;
  *out = e->timing;
  return STTS_OK;
}

uint64_t stts_launch_count(void) { return g_launch_count.load(std::memory_order_relaxed); }

float stts_last_vocoder_ms(const stts_engine* e, int which) {
  return (e && which >= 0 && which < 2) ? e->voc_ms[which] : -1.f;
}

int stts_timer_start(stts_engine* e) {
  if (!e) return STTS_ERR_INVALID;
  return guard_impl(e, [&] { CK(cudaEventRecord(e->ev[7], e->st)); });
}
int stts_timer_stop(stts_engine* e, float* ms) {
  if (!e || !ms) return STTS_ERR_INVALID;
  return guard_impl(e, [&] {
    CK(cudaEventRecord(e->ev_stop, e->st));
    CK(cudaEventSynchronize(e->ev_stop));
    CK(cudaEventElapsedTime(ms, e->ev[7], e->ev_stop));
  });
}

void* stts_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}
void stts_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

// ---------------------------------------------------------------- kernel-level test hooks
int stts_test_set_async(stts_engine* e, int on) {
  if (!e) return STTS_ERR_INVALID;
  e->test_async = on != 0;
  return STTS_OK;
}

int stts_test_gemm(stts_engine* e, int block_n, const void* a_bf16, int B, int T, int a_cols, int a_ld,
                   const void* w_bf16, int w_rows, int w_ld, int N, int K, int taps, int tap_shift0, int tap_step,
                   int groups, int a_group_koff, int w_group_rows, int out_group_cols, const float* bias, int act,
                   const int32_t* row_len, int rows_per_batch, int mask_bf16_only, const float* colscale,
                   const float* rowgate, int ld_gate, const float* residual, int ld_res, float* out_f32,
                   void* out_bf16, int ld_out) {
  if (!e) return STTS_ERR_INVALID;
  return guard_impl(e, [&] {
    GemmShape s;
    s.B = B; s.T = T; s.N = N; s.K = K; s.taps = taps; s.tap_shift0 = tap_shift0; s.tap_step = tap_step;
    s.groups = groups; s.a_group_koff = a_group_koff; s.w_group_rows = w_group_rows; s.out_group_cols = out_group_cols;
    s.res_group_cols = out_group_cols;
    GemmEpi ep;
    // act: GemmAct in the low 4 bits; +16 = GELU written as 2*gelu in fp16 (gelu2_f16); +32 = operands are fp16
    ep.bias = bias; ep.act = act & 15; ep.gelu2_f16 = (act >> 4) & 1; s.ab_f16 = (act >> 5) & 1; ep.row_len = row_len; ep.rows_per_batch = rows_per_batch; ep.mask_bf16_only = mask_bf16_only; ep.colscale = colscale;
    ep.rowgate = rowgate; ep.ld_gate = ld_gate; ep.residual = residual; ep.ld_res = ld_res; ep.out_f32 = out_f32;
    ep.out_bf16 = static_cast<bf16*>(out_bf16); ep.ld_out = ld_out;
    CK(launch_gemm(e->st, block_n, GemmA{static_cast<const bf16*>(a_bf16), a_cols, a_ld},
                   GemmW{static_cast<const bf16*>(w_bf16), w_rows, w_ld}, s, ep));
    if (!e->test_async) CK(cudaStreamSynchronize(e->st));
  });
}

int stts_test_gemm_split(stts_engine* e, int block_n, int splits, const void* a_bf16, int M, int K, const void* w_bf16, int N,
                         const float* bias, int gelu2_f16, const float* colscale, const float* residual, float* out_f32,
                         void* out_bf16) {
  if (!e) return STTS_ERR_INVALID;
  return guard_impl(e, [&] {
    GemmShape s;
    s.B = 1; s.T = M; s.N = N; s.K = K; s.splits = splits;
    GemmEpi ep;
    ep.bias = bias; ep.colscale = colscale; ep.residual = residual; ep.ld_res = N; ep.out_f32 = out_f32;
    ep.out_bf16 = static_cast<bf16*>(out_bf16); ep.ld_out = N;
    if (gelu2_f16) { ep.act = ACT_GELU; ep.gelu2_f16 = 1; }
    Tmp<float> scratch(e->st, static_cast<size_t>(gemm_split_scratch_floats(s, block_n)));
    Tmp<int> cnt(e->st, static_cast<size_t>(gemm_split_counters(s, block_n)));
    CK(cudaMemsetAsync(cnt.p, 0, static_cast<size_t>(gemm_split_counters(s, block_n)) * sizeof(int), e->st));
    ep.split_scratch = scratch; ep.split_counters = cnt;
    CK(launch_gemm(e->st, block_n, GemmA{static_cast<const bf16*>(a_bf16), K, K}, GemmW{static_cast<const bf16*>(w_bf16), N, K}, s, ep));
    // a second launch on the same counters: the kernel must have left them zero
    CK(launch_gemm(e->st, block_n, GemmA{static_cast<const bf16*>(a_bf16), K, K}, GemmW{static_cast<const bf16*>(w_bf16), N, K}, s, ep));
    if (e->test_async) {  // micro-benchmark mode: 20 more launches, timed on the stream
      cudaEvent_t t0, t1;
      CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
      CK(cudaEventRecord(t0, e->st));
      for (int i = 0; i < 20; ++i) {
        CK(launch_gemm(e->st, block_n, GemmA{static_cast<const bf16*>(a_bf16), K, K}, GemmW{static_cast<const bf16*>(w_bf16), N, K}, s, ep));
      }
      CK(cudaEventRecord(t1, e->st));
      CK(cudaEventSynchronize(t1));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, t0, t1));
      e->voc_ms[0] = ms / 20;  // read back through stts_last_vocoder_ms(e, 0)
      cudaEventDestroy(t0); cudaEventDestroy(t1);
    }
    CK(cudaStreamSynchronize(e->st));
  });
}

int stts_test_attention(stts_engine* e, const void* q, int B, int tq, int Hh, int hd, int hd_pad, const void* k0,
                        const void* v0, const int32_t* len0, int n0, const void* k1, const void* v1,
                        const int32_t* len1, int n1, const float* gate, int ld_gate, void* out) {
  if (!e) return STTS_ERR_INVALID;
  return guard_impl(e, [&] {
    AttnSeg segs[2];
    segs[0].k = static_cast<const bf16*>(k0); segs[0].v = static_cast<const bf16*>(v0); segs[0].len = len0; segs[0].n_max = n0;
    segs[1].k = static_cast<const bf16*>(k1); segs[1].v = static_cast<const bf16*>(v1); segs[1].len = len1; segs[1].n_max = n1;
    CK(attention_bf16(e->st, static_cast<const bf16*>(q), B, tq, Hh, hd, hd_pad, segs, k1 ? 2 : 1, gate, ld_gate, 0,
                      static_cast<bf16*>(out)));
    if (!e->test_async) CK(cudaStreamSynchronize(e->st));
  });
}

int stts_test_convnext_mix(stts_engine* e, const float* x, int B, int T, int C, const float* norm_w,
                           const float* conv_w, const float* conv_b, const float* gamma, const float* ffn_norm_w,
                           float* y, void* a_bf16) {
  if (!e) return STTS_ERR_INVALID;
  return guard_impl(e, [&] {
    CK(convnext_mix(e->st, x, B, T, C, norm_w, conv_w, conv_b, gamma, ffn_norm_w, 1e-5f, y, static_cast<bf16*>(a_bf16)));
    if (!e->test_async) CK(cudaStreamSynchronize(e->st));
  });
}

int stts_test_ffn_fused(stts_engine* e, const void* a_bf16, const float* y, long long M, int C, const void* w1_bf16,
                        const float* b1, const void* w2_f16, const float* b2, const float* ffn_gamma, float* out,
                        void* out_bf16) {
  if (!e) return STTS_ERR_INVALID;
  return guard_impl(e, [&] {
    CK(ffn_fused(e->st, static_cast<const bf16*>(a_bf16), y, M, C, static_cast<const bf16*>(w1_bf16), b1, w2_f16, b2,
                 ffn_gamma, out, static_cast<bf16*>(out_bf16)));
    if (!e->test_async) CK(cudaStreamSynchronize(e->st));
  });
}

int stts_test_convnext_fused(stts_engine* e, const float* x, int B, int T, int C, const float* norm_w,
                             const float* conv_w, const float* conv_b, const float* gamma, const float* ffn_norm_w,
                             const void* w1_bf16, const float* b1, const void* w2_f16, const float* b2,
                             const float* ffn_gamma, float* out, void* out_bf16) {
  if (!e) return STTS_ERR_INVALID;
  return guard_impl(e, [&] {
    CK(convnext_fused(e->st, x, B, T, C, norm_w, conv_w, conv_b, gamma, ffn_norm_w, static_cast<const bf16*>(w1_bf16), b1,
                      w2_f16, b2, ffn_gamma, 1e-5f, out, static_cast<bf16*>(out_bf16)));
    if (!e->test_async) CK(cudaStreamSynchronize(e->st));
  });
}


namespace {
ChainWeights chain_weights_of(const stts_test_chain_args* a) {
  ChainWeights w;
  w.wqkvg = static_cast<const bf16*>(a->wqkvg); w.wo = static_cast<const bf16*>(a->wo);
  w.w13 = static_cast<const bf16*>(a->w13); w.w2 = static_cast<const bf16*>(a->w2); w.wvel = static_cast<const bf16*>(a->wvel);
  w.bqkvg = a->bqkvg; w.b13 = a->b13; w.b2 = a->b2; w.bvel = a->bvel; w.qn = a->qn; w.kn = a->kn;
  w.cos_t = a->cos_t; w.sin_t = a->sin_t;
  return w;
}
}  // namespace

int stts_test_chain(stts_engine* e, const stts_test_chain_args* a) {
  if (!e || !a) return STTS_ERR_INVALID;
  return guard_impl(e, [&] {
    ChainBuffers b;
    b.x = a->x; b.xb = static_cast<bf16*>(a->xb); b.stats = a->stats; b.qkv = static_cast<bf16*>(a->qkv); b.gate = a->gate;
    b.ob = static_cast<bf16*>(a->ob); b.hb = static_cast<bf16*>(a->hb); b.vel = a->vel; b.ready = a->ready;
    b.trace = reinterpret_cast<unsigned long long*>(a->trace);
    ChainCall c;
    c.M = a->M; c.T = a->T; c.frames = a->frames; c.mod = a->mod; c.fold = a->fold; c.n_phases = a->n_phases;
    if (a->n_phases < 0 || a->n_phases > kChainMaxPhases) throw Err(STTS_ERR_INVALID, "n_phases out of range");
    for (int i = 0; i < a->n_phases; ++i) { c.kind[i] = a->kind[i]; c.blk[i] = a->blk[i]; }
    c.qkv_db = a->qkv_db != 0;
    c.attn.B = a->B; c.attn.R = a->R; c.attn.P = a->P; c.attn.ref_len = a->ref_len; c.attn.ph_len = a->ph_len;
    c.attn.kv_ref = static_cast<const bf16*>(a->kv_ref); c.attn.kv_text = static_cast<const bf16*>(a->kv_text);
    CK(launch_dit_chain(e->st, chain_weights_of(a), b, c));
    if (!e->test_async) CK(cudaStreamSynchronize(e->st));
  });
}

int stts_test_chain_fold(stts_engine* e, const stts_test_chain_args* a, float* fold_out) {
  if (!e || !a || !fold_out) return STTS_ERR_INVALID;
  return guard_impl(e, [&] {
    CK(chain_fold_table(e->st, chain_weights_of(a), a->mod, fold_out));
    if (!e->test_async) CK(cudaStreamSynchronize(e->st));
  });
}

int64_t stts_test_chain_fold_floats(void) { return kFoldFloats; }

int stts_test_chain_stats_cast(stts_engine* e, const float* x, int M, const float* scale, void* xb, float* stats) {
  if (!e) return STTS_ERR_INVALID;
  return guard_impl(e, [&] {
    CK(chain_stats_cast(e->st, x, M, scale, static_cast<bf16*>(xb), stats));
    if (!e->test_async) CK(cudaStreamSynchronize(e->st));
  });
}

}  // extern "C"
