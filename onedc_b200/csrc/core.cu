// Library plumbing: error string, device info, launch counter.
#include <stdarg.h>

#include <atomic>

#include "../../include/onedc_b200.h"
#include "common.cuh"

namespace onedc {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace onedc

extern "C" const char* onedc_last_error(void) { return onedc::g_err; }
extern "C" int onedc_version(void) { return 100; }
extern "C" int64_t onedc_launch_count(int reset) {
  long long v = onedc::g_launches.load();
  if (reset) onedc::g_launches.store(0);
  return v;
}
