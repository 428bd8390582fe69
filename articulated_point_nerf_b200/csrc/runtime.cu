// Error state, launch counter and version of libapn_sm100.so.
#include <stdarg.h>

#include <atomic>

#include <nvtx3/nvToolsExt.h>

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

// Two library-owned side streams for fork/join concurrency INSIDE one entry point (small independent GEMMs that each
// fill less than half the GPU).  A call forks from the caller's stream with an event, runs the independent work on the
// side streams, and joins back into the caller's stream before it returns, so callers keep plain single-stream
// semantics (and CUDA-graph capture of the caller's stream captures the fork/join too).
static ApnSide g_side;
static bool g_side_ready = false;
int apn_side_streams(ApnSide** out) {
  if (!g_side_ready) {
    for (int i = 0; i < 2; ++i) {
      APN_CUDA(cudaStreamCreateWithFlags(&g_side.s[i], cudaStreamNonBlocking));
      APN_CUDA(cudaEventCreateWithFlags(&g_side.join[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 3; ++i) APN_CUDA(cudaEventCreateWithFlags(&g_side.fork[i], cudaEventDisableTiming));
    g_side_ready = true;
  }
  *out = &g_side;
  return 0;
}

// NVTX ranges around the C-ABI call groups, named after the reference's profiler ranges (torch.profiler.record_function at
// lib/temporalpoints.py:421-653, lib/pointwarper.py:217-241) so that a timeline of this path reads like one of the reference.
// nvtx3 is header-only: without an attached tool the calls are a branch on a null function table.
extern "C" int apn_range_push(const char* name) { return nvtxRangePushA(name ? name : ""); }
extern "C" int apn_range_pop(void) { return nvtxRangePop(); }
