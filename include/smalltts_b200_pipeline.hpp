// smalltts_b200_pipeline.hpp -- header-only C++17 host side above the C ABI (include/smalltts_b200.h) for COMPILED callers.
//
// The reference's compiled caller of the hot path is its Rust server: `Pipeline` in src/server/src/pipeline.rs:40-112
// (load four onnxruntime sessions; `synthesize` / `synthesize_timed`: codec encode -> condition encode -> 4 denoiser
// steps -> codec decode, with the `Timing` split of :29-37) and the table-printing benchmark src/server/src/bin/bench.rs.
// cargo is not part of this image, so the mirror of that interface is written in C++ (same names, argument meaning and
// error behaviour: every failure is an exception carrying stts_last_error, like the reference's anyhow::Result).
//
//   stts::Pipeline p = stts::Pipeline::load("dit.sttsw", "decoder.sttsw", "encoder.sttsw");   // tools/convert_weights.py
//   auto [audio, t] = p.synthesize_timed(ref_audio_24k, token_ids, 10.0f);
//   auto many      = p.synthesize_many({refA, refB}, {tokA, tokB}, {2.0f, 10.0f});             // one ragged engine pass
//
// Weight files are `.sttsw` containers (smalltts_b200/weights.py: magic "STTSW001", u64 index length, JSON index
// {name: {shape, dtype, offset, nbytes}}, 64-byte aligned little-endian payloads, fp32 or bf16).
#ifndef SMALLTTS_B200_PIPELINE_HPP_
#define SMALLTTS_B200_PIPELINE_HPP_

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <random>
#include <string>
#include <utility>
#include <vector>

#include "smalltts_b200.h"

namespace stts {

constexpr float SR = 24000.0f;  // pipeline.rs:11
constexpr float HOP = 3200.0f;  // pipeline.rs:12
constexpr int STEPS = 4;        // pipeline.rs:14

struct Timing {  // pipeline.rs:29-37, milliseconds
  double codec_enc_ms = 0, cond_enc_ms = 0, denoise_ms = 0, codec_dec_ms = 0, total_ms = 0;
};

struct Error : std::runtime_error {
  int status;
  Error(int s, const std::string& m) : std::runtime_error(m), status(s) {}
};

// ---------------------------------------------------------------------------------------------- .sttsw reader
struct PackedTensor {
  std::string name;
  std::vector<int64_t> shape;
  std::vector<float> data;  // always fp32 here (bf16 payloads are widened)
};

namespace detail {
inline void skip_ws(const std::string& s, size_t& i) {
  while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r' || s[i] == ',')) ++i;
}
inline std::string parse_string(const std::string& s, size_t& i) {
  if (i >= s.size() || s[i] != '"') throw Error(STTS_ERR_WEIGHTS, "sttsw index: expected a string");
  std::string out;
  for (++i; i < s.size() && s[i] != '"'; ++i) {
    if (s[i] == '\\' && i + 1 < s.size()) ++i;  // state-dict names never need more than this
    out.push_back(s[i]);
  }
  if (i >= s.size()) throw Error(STTS_ERR_WEIGHTS, "sttsw index: unterminated string");
  ++i;
  return out;
}
inline int64_t parse_int(const std::string& s, size_t& i) {
  size_t j = i;
  while (j < s.size() && (s[j] == '-' || (s[j] >= '0' && s[j] <= '9'))) ++j;
  if (j == i) throw Error(STTS_ERR_WEIGHTS, "sttsw index: expected a number");
  const int64_t v = std::stoll(s.substr(i, j - i));
  i = j;
  return v;
}
inline void expect(const std::string& s, size_t& i, char c) {
  skip_ws(s, i);
  if (i >= s.size() || s[i] != c) throw Error(STTS_ERR_WEIGHTS, std::string("sttsw index: expected '") + c + "'");
  ++i;
}
}  // namespace detail

inline std::vector<PackedTensor> read_sttsw(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw Error(STTS_ERR_WEIGHTS, path + ": cannot open");
  char magic[8];
  uint64_t n = 0;
  f.read(magic, 8);
  f.read(reinterpret_cast<char*>(&n), 8);
  if (!f || std::memcmp(magic, "STTSW001", 8) != 0 || n > (1ull << 30)) throw Error(STTS_ERR_WEIGHTS, path + ": not a .sttsw file");
  std::string idx(n, '\0');
  f.read(&idx[0], static_cast<std::streamsize>(n));
  if (!f) throw Error(STTS_ERR_WEIGHTS, path + ": truncated index");
  const uint64_t base = 16 + n;
  std::vector<PackedTensor> out;
  size_t i = 0;
  detail::expect(idx, i, '{');
  for (;;) {
    detail::skip_ws(idx, i);
    if (i < idx.size() && idx[i] == '}') break;
    PackedTensor t;
    t.name = detail::parse_string(idx, i);
    detail::expect(idx, i, ':');
    detail::expect(idx, i, '{');
    std::string dtype = "float32";
    int64_t offset = -1, nbytes = -1;
    for (;;) {
      detail::skip_ws(idx, i);
      if (i < idx.size() && idx[i] == '}') {
        ++i;
        break;
      }
      const std::string key = detail::parse_string(idx, i);
      detail::expect(idx, i, ':');
      detail::skip_ws(idx, i);
      if (key == "shape") {
        detail::expect(idx, i, '[');
        for (;;) {
          detail::skip_ws(idx, i);
          if (i < idx.size() && idx[i] == ']') {
            ++i;
            break;
          }
          t.shape.push_back(detail::parse_int(idx, i));
        }
      } else if (key == "dtype") {
        dtype = detail::parse_string(idx, i);
      } else if (key == "offset") {
        offset = detail::parse_int(idx, i);
      } else if (key == "nbytes") {
        nbytes = detail::parse_int(idx, i);
      } else {
        throw Error(STTS_ERR_WEIGHTS, path + ": unknown index field " + key);
      }
    }
    int64_t numel = 1;
    for (int64_t d : t.shape) numel *= d;
    const bool bf16 = dtype == "bfloat16";
    if (offset < 0 || nbytes != numel * (bf16 ? 2 : 4) || (!bf16 && dtype != "float32")) {
      throw Error(STTS_ERR_WEIGHTS, path + ": bad index entry for " + t.name);
    }
    t.data.resize(static_cast<size_t>(numel));
    f.seekg(static_cast<std::streamoff>(base + static_cast<uint64_t>(offset)));
    if (bf16) {
      std::vector<uint16_t> raw(static_cast<size_t>(numel));
      f.read(reinterpret_cast<char*>(raw.data()), nbytes);
      for (int64_t k = 0; k < numel; ++k) {
        const uint32_t u = static_cast<uint32_t>(raw[static_cast<size_t>(k)]) << 16;
        std::memcpy(&t.data[static_cast<size_t>(k)], &u, 4);
      }
    } else {
      f.read(reinterpret_cast<char*>(t.data.data()), nbytes);
    }
    if (!f) throw Error(STTS_ERR_WEIGHTS, path + ": truncated payload of " + t.name);
    out.push_back(std::move(t));
  }
  return out;
}

// ---------------------------------------------------------------------------------------------- Pipeline
class Pipeline {
 public:
  Pipeline(const Pipeline&) = delete;
  Pipeline& operator=(const Pipeline&) = delete;
  Pipeline(Pipeline&& o) noexcept : e_(o.e_), base_seed_(o.base_seed_), calls_(o.calls_) { o.e_ = nullptr; }
  ~Pipeline() {
    if (e_) stts_destroy(e_);
  }

  // pipeline.rs:40-48 `Pipeline::load`: the four models (cond encoder + denoiser are one DiT checkpoint here).
  // `encoder` may be empty: then `synthesize*` need reference LATENTS (synthesize_latents) instead of audio.
  static Pipeline load(const std::string& dit, const std::string& decoder, const std::string& encoder = "", int device = 0) {
    stts_config cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.device = device;
    stts_engine* e = nullptr;
    const int rc = stts_create(&cfg, &e);
    if (rc != STTS_OK) throw Error(rc, std::string("stts_create: ") + stts_last_error(nullptr));
    Pipeline p(e);
    const std::string* files[3] = {&dit, &decoder, &encoder};
    for (int model = 0; model < 3; ++model) {
      if (files[model]->empty()) continue;
      for (const PackedTensor& t : read_sttsw(*files[model])) {
        p.check(stts_load_weight(e, model, t.name.c_str(), t.data.data(), static_cast<int>(t.shape.size()),
                                 t.shape.empty() ? nullptr : t.shape.data()),
                "stts_load_weight");
      }
    }
    p.check(stts_finalize_weights(e), "stts_finalize_weights");
    return p;
  }

  // pipeline.rs:50-58
  std::vector<float> synthesize(const std::vector<float>& ref_audio, const std::vector<int64_t>& token_ids, float duration_sec) {
    return synthesize_timed(ref_audio, token_ids, duration_sec).first;
  }

  // pipeline.rs:60-112: one request; seq_len = ceil(duration * SR / HOP).max(1)
  std::pair<std::vector<float>, Timing> synthesize_timed(const std::vector<float>& ref_audio,
                                                         const std::vector<int64_t>& token_ids, float duration_sec) {
    Timing t;
    auto out = synthesize_many({ref_audio}, {token_ids}, {duration_sec}, &t);
    return {std::move(out[0]), t};
  }

  // Several requests as ONE ragged engine pass (the reference loops over them, bench.rs:44-47): reference clips are
  // right-padded (the codec encoder is causal: a clip's latents do not depend on the padding), prompts are masked by length.
  std::vector<std::vector<float>> synthesize_many(const std::vector<std::vector<float>>& ref_audio,
                                                  const std::vector<std::vector<int64_t>>& token_ids,
                                                  const std::vector<float>& duration_sec, Timing* timing = nullptr,
                                                  uint64_t seed = kAutoSeed) {
    const int B = static_cast<int>(ref_audio.size());
    if (B == 0 || token_ids.size() != ref_audio.size() || duration_sec.size() != ref_audio.size()) {
      throw Error(STTS_ERR_INVALID, "ref_audio, token_ids and duration_sec must be equally long and non-empty");
    }
    const int hop = static_cast<int>(HOP);
    std::vector<int64_t> ref_len(B), ph_len(B), frames(B);
    int R = 0, P = 1, T = 0;
    for (int b = 0; b < B; ++b) {
      ref_len[b] = static_cast<int64_t>(ref_audio[b].size()) / hop;
      if (ref_len[b] < 1) throw Error(STTS_ERR_INVALID, "reference audio shorter than one codec hop (3200 samples at 24 kHz)");
      ph_len[b] = static_cast<int64_t>(token_ids[b].size());
      const float fr = std::ceil(duration_sec[b] * SR / HOP);
      frames[b] = fr < 1.0f ? 1 : static_cast<int64_t>(fr);
      R = std::max<int>(R, static_cast<int>(ref_len[b]));
      P = std::max<int>(P, static_cast<int>(ph_len[b]));
      T = std::max<int>(T, static_cast<int>(frames[b]));
    }
    std::vector<float> audio_in(static_cast<size_t>(B) * R * hop, 0.0f);
    for (int b = 0; b < B; ++b) {
      std::memcpy(&audio_in[static_cast<size_t>(b) * R * hop], ref_audio[b].data(), static_cast<size_t>(ref_len[b]) * hop * 4);
    }
    std::vector<float> latents(static_cast<size_t>(B) * R * STTS_LATENT_DIM);
    check(stts_encode_audio(e_, audio_in.data(), B, R * hop, STTS_MEM_HOST, latents.data()), "stts_encode_audio");
    stts_timing tm_enc;
    check(stts_get_timings(e_, &tm_enc), "stts_get_timings");
    return run(latents, ref_len, token_ids, ph_len, frames, B, R, P, T, seed, tm_enc.codec_enc_ms, timing);
  }

  // The Python API's entry: reference LATENTS [R, 64] instead of audio (infer/onnx.py:68-83).
  std::vector<float> synthesize_latents(const std::vector<float>& ref_latents, const std::vector<int64_t>& token_ids,
                                        float duration_sec, Timing* timing = nullptr, uint64_t seed = kAutoSeed) {
    const int R = static_cast<int>(ref_latents.size() / STTS_LATENT_DIM);
    if (R < 1 || ref_latents.size() % STTS_LATENT_DIM != 0) throw Error(STTS_ERR_INVALID, "ref_latents must be [R, 64]");
    const float fr = std::ceil(duration_sec * SR / HOP);
    std::vector<int64_t> ref_len{R}, ph_len{static_cast<int64_t>(token_ids.size())}, frames{fr < 1.0f ? 1 : static_cast<int64_t>(fr)};
    const int P = std::max<int>(1, static_cast<int>(token_ids.size()));
    return std::move(run(ref_latents, ref_len, {token_ids}, ph_len, frames, 1, R, P, static_cast<int>(frames[0]), seed, 0.0f, timing)[0]);
  }

  stts_engine* handle() { return e_; }

  // Noise: the reference draws fresh noise for every request (pipeline.rs:249-255).  Without an explicit seed every call
  // gets base_seed + call counter; the base comes from std::random_device, so two pipelines (replicas, restarts) do
  // not replay each other's streams.
  static constexpr uint64_t kAutoSeed = ~static_cast<uint64_t>(0);
  void set_base_seed(uint64_t s) { base_seed_ = s; calls_ = 0; }

 private:
  explicit Pipeline(stts_engine* e) : e_(e) {
    std::random_device rd;
    base_seed_ = (static_cast<uint64_t>(rd()) << 32) ^ rd();
  }

  void check(int rc, const char* what) const {
    if (rc != STTS_OK) throw Error(rc, std::string(what) + ": " + stts_last_error(e_));
  }

  std::vector<std::vector<float>> run(const std::vector<float>& latents, const std::vector<int64_t>& ref_len,
                                      const std::vector<std::vector<int64_t>>& token_ids, const std::vector<int64_t>& ph_len,
                                      const std::vector<int64_t>& frames, int B, int R, int P, int T, uint64_t seed,
                                      float codec_enc_ms, Timing* timing) {
    std::vector<int64_t> ids(static_cast<size_t>(B) * P, 0);
    for (int b = 0; b < B; ++b) {
      for (size_t k = 0; k < token_ids[b].size(); ++k) ids[static_cast<size_t>(b) * P + k] = token_ids[b][k];
    }
    const int hop = static_cast<int>(HOP);
    std::vector<float> audio(static_cast<size_t>(B) * T * hop);
    if (seed == kAutoSeed) seed = base_seed_ + calls_++;
    check(stts_synthesize(e_, latents.data(), ref_len.data(), ids.data(), ph_len.data(), frames.data(), B, R, P, T, STEPS,
                          nullptr, nullptr, seed, STTS_MEM_HOST, audio.data()),
          "stts_synthesize");
    if (timing) {
      stts_timing tm;
      check(stts_get_timings(e_, &tm), "stts_get_timings");
      timing->codec_enc_ms = codec_enc_ms;
      timing->cond_enc_ms = tm.cond_enc_ms;
      timing->denoise_ms = tm.denoise_ms;
      timing->codec_dec_ms = tm.codec_dec_ms;
      timing->total_ms = codec_enc_ms + tm.total_ms;
    }
    std::vector<std::vector<float>> out(static_cast<size_t>(B));
    for (int b = 0; b < B; ++b) {
      const float* src = &audio[static_cast<size_t>(b) * T * hop];
      out[static_cast<size_t>(b)].assign(src, src + static_cast<size_t>(frames[b]) * hop);
    }
    return out;
  }

  stts_engine* e_ = nullptr;
  uint64_t base_seed_ = 0, calls_ = 0;
};

}  // namespace stts

#endif  // SMALLTTS_B200_PIPELINE_HPP_
