// Library-wide ABI helpers: error text, version, launch counter.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include <stdlib.h>
#include "common.cuh"

namespace pc {
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
// PC_PDL=1 / 0 forces programmatic dependent launch on / off; by default it is used for eager launches only: measured on
// the cnn_deep step, eager 5.56 -> 5.49 ms with PDL, but inside the captured whole-step graph 6.51 -> 6.60 ms (the early-
// resident CTAs take SM slots the overlapped weight-gradient lane would otherwise fill).
bool pdl_enabled(cudaStream_t stream) {
  static const int mode = [] {
    const char* e = getenv("PC_PDL");
    return e == nullptr ? -1 : (e[0] == '1' ? 1 : 0);
  }();
  if (mode >= 0) return mode == 1;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) return false;
  return st == cudaStreamCaptureStatusNone;
}
}  // namespace pc

extern "C" const char* pc_last_error(void) { return pc::g_err; }
extern "C" int pc_abi_version(void) { return 1; }
extern "C" unsigned long long pc_launch_count(void) { return pc::g_launches.load(std::memory_order_relaxed); }
