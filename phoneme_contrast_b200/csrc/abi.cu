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
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("PC_PDL");   // opt-in: measured slightly slower than plain launches inside the captured step
    return e && e[0] == '1';
  }();
  return on;
}
}  // namespace pc

extern "C" const char* pc_last_error(void) { return pc::g_err; }
extern "C" int pc_abi_version(void) { return 1; }
extern "C" unsigned long long pc_launch_count(void) { return pc::g_launches.load(std::memory_order_relaxed); }
