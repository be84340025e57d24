// Library plumbing: error string, device info, launch counter.
#include <stdarg.h>
#include <stdlib.h>

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

static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("ONEDC_PDL");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;      // opt-in: measured slower inside graphs on B200
    g_pdl.store(v);
  }
  return v != 0;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace onedc

extern "C" const char* onedc_last_error(void) { return onedc::g_err; }
extern "C" int onedc_set_pdl(int on) {
  const int old = onedc::pdl_enabled() ? 1 : 0;
  onedc::g_pdl.store(on ? 1 : 0);
  return old;
}
extern "C" int onedc_version(void) { return 100; }
extern "C" int64_t onedc_launch_count(int reset) {
  long long v = onedc::g_launches.load();
  if (reset) onedc::g_launches.store(0);
  return v;
}
