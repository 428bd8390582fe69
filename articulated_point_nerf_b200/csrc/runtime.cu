// Error state, launch counter and version of libapn_sm100.so.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

static thread_local char g_err[1024] = "";
static std::atomic<unsigned long long> g_launches{0};

void apn_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void apn_count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

extern "C" int apn_version(void) { return 100; }
extern "C" const char* apn_last_error(void) { return g_err; }
extern "C" unsigned long long apn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
