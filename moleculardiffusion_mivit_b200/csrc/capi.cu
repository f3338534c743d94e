// Library-wide pieces of the C ABI: version, error text, launch counter.
#include <stdarg.h>
#include <atomic>

#include "common.cuh"
#include "../../include/mivit.h"

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void mivit_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void mivit_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" int mivit_abi_version(void) { return MIVIT_ABI_VERSION; }
extern "C" const char* mivit_last_error(void) { return g_err; }
extern "C" int64_t mivit_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" void mivit_reset_launch_count(void) { g_launches.store(0); }
