// Library plumbing: thread-local last-error string, ABI version, launch counter.
#include <stdarg.h>
#include <stdlib.h>
#include <atomic>
#include "common.cuh"
#include "mrcnn_b200.h"

static thread_local char g_err[1024] = "";
std::atomic<unsigned long long> g_mrcnn_launches{0};

void mrcnn_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void mrcnn_count_launch(unsigned long long n) { g_mrcnn_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" const char* mrcnn_last_error(void) { return g_err; }
extern "C" int mrcnn_abi_version(void) { return 2; }
extern "C" unsigned long long mrcnn_kernel_launch_count(void) { return g_mrcnn_launches.load(); }

// programmatic dependent launch is on unless MRCNN_B200_PDL=0 (read once)
bool mrcnn_pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("MRCNN_B200_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}
