// The reference's server benchmark (src/server/src/bin/bench.rs) on the B200 engine, through the C++ mirror of its
// Pipeline (include/smalltts_b200_pipeline.hpp): 440 Hz sine reference of 2 s, tokens 1..30, durations 2 / 5 / 10 s,
// batches 1 / 2 / 4 / 8 -- once the reference's way (a sequential loop of single requests, bench.rs:44-47) and once as a
// single ragged engine pass per batch.
//
//   g++ -O2 -std=c++17 -Iinclude examples/bench_pipeline.cpp -Lsmalltts_b200 -lsmalltts_b200
//       -Wl,-rpath,$PWD/smalltts_b200 -o bench_pipeline
//   python tools/make_synthetic_sttsw.py /tmp/w            # or tools/convert_weights.py on real assets
//   ./bench_pipeline /tmp/w/dit.sttsw /tmp/w/decoder.sttsw /tmp/w/encoder.sttsw
#include <chrono>
#include <cmath>
#include <cstdio>
#include <vector>

#include "smalltts_b200_pipeline.hpp"

static std::vector<float> sine_wave(float duration_sec, int sample_rate) {  // bench.rs:8-13
  const int n = static_cast<int>(duration_sec * sample_rate);
  std::vector<float> w(n);
  for (int i = 0; i < n; ++i) w[i] = std::sin(2.0f * 3.14159265358979f * 440.0f * i / sample_rate);
  return w;
}

int main(int argc, char** argv) {
  if (argc < 4) {
    std::fprintf(stderr, "usage: %s dit.sttsw decoder.sttsw encoder.sttsw [device]\n", argv[0]);
    return 2;
  }
  constexpr int kWarmup = 1, kRuns = 3;  // bench.rs:3-4
  try {
    std::printf("smalltts_b200 pipeline benchmark\n\n");
    stts::Pipeline pipe = stts::Pipeline::load(argv[1], argv[2], argv[3], argc > 4 ? std::atoi(argv[4]) : 0);
    std::printf("pipeline loaded\n");
    const std::vector<float> ref_audio = sine_wave(2.0f, 24000);
    std::vector<int64_t> tokens(30);
    for (int i = 0; i < 30; ++i) tokens[i] = i + 1;
    const float durations[3] = {2.0f, 5.0f, 10.0f};
    const int batches[4] = {1, 2, 4, 8};
    for (int batch : batches) {
      std::printf("\nbatch = %d\n", batch);
      std::printf("  %6s  %9s  %9s  %9s  %9s  %10s  %8s  | %12s  %8s\n", "dur(s)", "codec_enc", "cond_enc", "denoise",
                  "codec_dec", "total(ms)", "RTF", "1 pass (ms)", "RTF");
      for (float dur : durations) {
        stts::Timing sum, t;
        double seq_ms = 0, one_ms = 0;
        for (int run = 0; run < kWarmup + kRuns; ++run) {
          // sequential, like the reference
          stts::Timing acc;
          const auto w0 = std::chrono::steady_clock::now();
          for (int b = 0; b < batch; ++b) {
            t = pipe.synthesize_timed(ref_audio, tokens, dur).second;
            acc.codec_enc_ms += t.codec_enc_ms;
            acc.cond_enc_ms += t.cond_enc_ms;
            acc.denoise_ms += t.denoise_ms;
            acc.codec_dec_ms += t.codec_dec_ms;
          }
          const auto w1 = std::chrono::steady_clock::now();
          // the same requests as one ragged pass
          std::vector<std::vector<float>> refs(batch, ref_audio);
          std::vector<std::vector<int64_t>> toks(batch, tokens);
          std::vector<float> durs(batch, dur);
          pipe.synthesize_many(refs, toks, durs);
          const auto w2 = std::chrono::steady_clock::now();
          if (run >= kWarmup) {
            sum.codec_enc_ms += acc.codec_enc_ms;
            sum.cond_enc_ms += acc.cond_enc_ms;
            sum.denoise_ms += acc.denoise_ms;
            sum.codec_dec_ms += acc.codec_dec_ms;
            seq_ms += std::chrono::duration<double, std::milli>(w1 - w0).count();
            one_ms += std::chrono::duration<double, std::milli>(w2 - w1).count();
          }
        }
        const double n = kRuns * static_cast<double>(batch);
        const double audio_sec = std::ceil(dur * stts::SR / stts::HOP) * stts::HOP / stts::SR;  // bench.rs:66-67
        std::printf("  %6.1f  %9.2f  %9.2f  %9.2f  %9.2f  %10.2f  %8.5f  | %12.2f  %8.5f\n", dur, sum.codec_enc_ms / n,
                    sum.cond_enc_ms / n, sum.denoise_ms / n, sum.codec_dec_ms / n, seq_ms / kRuns,
                    seq_ms / kRuns / 1000.0 / (audio_sec * batch), one_ms / kRuns, one_ms / kRuns / 1000.0 / (audio_sec * batch));
      }
    }
  } catch (const stts::Error& e) {
    std::fprintf(stderr, "error (status %d): %s\n", e.status, e.what());
    return 1;
  }
  return 0;
}
