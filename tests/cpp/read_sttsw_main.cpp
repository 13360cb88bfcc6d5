// Test helper: dump a .sttsw container as read by the C++ host side (include/smalltts_b200_pipeline.hpp) as JSON lines
// {"name": ..., "shape": [...], "sum": ..., "first": ..., "last": ...} for tests/test_cpp_host_cpu.py.
#include <cstdio>

#include "smalltts_b200_pipeline.hpp"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  try {
    for (const stts::PackedTensor& t : stts::read_sttsw(argv[1])) {
      double sum = 0;
      for (float v : t.data) sum += v;
      std::printf("{\"name\": \"%s\", \"shape\": [", t.name.c_str());
      for (size_t i = 0; i < t.shape.size(); ++i) std::printf("%s%lld", i ? ", " : "", static_cast<long long>(t.shape[i]));
      std::printf("], \"sum\": %.9g, \"first\": %.9g, \"last\": %.9g}\n", sum, t.data.empty() ? 0.0 : t.data.front(),
                  t.data.empty() ? 0.0 : t.data.back());
    }
  } catch (const stts::Error& e) {
    std::fprintf(stderr, "error (status %d): %s\n", e.status, e.what());
    return 1;
  }
  return 0;
}
