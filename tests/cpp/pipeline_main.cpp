// Test driver for include/smalltts_b200_pipeline.hpp (tests/test_gpu_cpp_host.py): load .sttsw files, run one ragged
// two-request pass with an explicit seed plus one timed single request, write the waveforms as raw fp32.
#include <cmath>
#include <cstdio>
#include <vector>

#include "smalltts_b200_pipeline.hpp"

static std::vector<float> tone(float seconds, float hz) {
  std::vector<float> w(static_cast<size_t>(seconds * 24000));
  for (size_t i = 0; i < w.size(); ++i) w[i] = 0.3f * std::sin(2.0f * 3.14159265358979f * hz * i / 24000.0f);
  return w;
}

int main(int argc, char** argv) {
  if (argc < 5) {
    std::fprintf(stderr, "usage: %s dit.sttsw decoder.sttsw encoder.sttsw out.f32\n", argv[0]);
    return 2;
  }
  try {
    stts::Pipeline pipe = stts::Pipeline::load(argv[1], argv[2], argv[3]);
    const std::vector<std::vector<float>> refs = {tone(2.0f, 440.0f), tone(1.2f, 220.0f)};
    const std::vector<std::vector<int64_t>> toks = {{5, 9, 20, 33, 7}, {101, 3, 44}};
    stts::Timing tm;
    auto many = pipe.synthesize_many(refs, toks, {1.01f, 0.5f}, &tm, /*seed=*/1234);
    auto one = pipe.synthesize_timed(refs[0], toks[0], 2.0f);
    auto again = pipe.synthesize_timed(refs[0], toks[0], 2.0f);  // fresh noise per request (pipeline.rs:249-255)
    FILE* f = std::fopen(argv[4], "wb");
    if (!f) return 3;
    for (const auto& a : many) std::fwrite(a.data(), 4, a.size(), f);
    std::fclose(f);
    double diff = 0;
    for (size_t i = 0; i < one.first.size(); ++i) diff += std::fabs(one.first[i] - again.first[i]);
    std::printf("{\"n0\": %zu, \"n1\": %zu, \"n_single\": %zu, \"batch_total_ms\": %.4f, \"denoise_ms\": %.4f, "
                "\"codec_enc_ms\": %.4f, \"single_total_ms\": %.4f, \"repeat_abs_diff\": %.6f}\n",
                many[0].size(), many[1].size(), one.first.size(), tm.total_ms, tm.denoise_ms, tm.codec_enc_ms,
                one.second.total_ms, diff);
  } catch (const stts::Error& e) {
    std::fprintf(stderr, "error (status %d): %s\n", e.status, e.what());
    return 1;
  }
  return 0;
}
