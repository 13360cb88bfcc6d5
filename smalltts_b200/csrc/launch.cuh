// Kernel launch helper: every kernel of the engine is launched with the programmatic-stream-serialization attribute
// (Programmatic Dependent Launch).  Contract for kernels: call ptx::pdl_wait() before the first access to global
// memory that a predecessor may have written or may still read, and never exit without having called it (so that
// completion stays transitive along the stream); ptx::pdl_trigger() lets the successor's CTAs be scheduled early, so
// its launch latency and prologue overlap this kernel's tail.  On by default (A/B on B200 at BASELINE configs[1] under
// graph replay: 11.45 -> 10.77 ms per step); STTS_PDL=0 turns the attribute off.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <utility>

namespace stts {

// Function attributes (cudaFuncAttributeMaxDynamicSharedMemorySize) belong to the current device: a process that drives
// several GPUs (SmallTTS(devices=[...])) has to set them once per device, and engine replicas launch from several host
// threads.  `static PerDeviceOnce once; once.run([&] { return cudaFuncSetAttribute(...); })` does both.
class PerDeviceOnce {
 public:
  template <typename F>
  cudaError_t run(F&& fn) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done_.load(std::memory_order_acquire) & bit) return cudaSuccess;
    std::lock_guard<std::mutex> g(m_);
    if (done_.load(std::memory_order_relaxed) & bit) return cudaSuccess;
    e = fn();
    if (e == cudaSuccess) done_.fetch_or(bit, std::memory_order_release);
    return e;
  }

 private:
  std::mutex m_;
  std::atomic<unsigned long long> done_{0};
};

// Kernels launched (really executed) by this library since process start: bench.py's gpu_launches.  Engines may be driven
// from several host threads (replicas, SmallTTS(devices=[...])), hence atomic.  Launches recorded into a CUDA graph
// during stream capture do not run, so they are not counted here; a graph replay adds the number of kernel nodes the
// graph holds (engine.cu: capture()).
extern std::atomic<unsigned long long> g_launch_count;
extern thread_local int tl_capturing;  // > 0 while this host thread records a graph
inline void count_launch() {
  if (tl_capturing == 0) g_launch_count.fetch_add(1, std::memory_order_relaxed);
}

inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("STTS_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}

// Same, as clusters of two CTAs along x (CTA pairs for cta_group::2 kernels); grid.x must be even.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k_pair(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}

}  // namespace stts
