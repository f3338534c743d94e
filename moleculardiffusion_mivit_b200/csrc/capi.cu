// Library-wide pieces of the C ABI: version, error text, launch counter.
#include <stdarg.h>
#include <atomic>

#include "common.cuh"
#include "tma.cuh"
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

mivit_tensor_map_encode_fn mivit_tensor_map_encoder() {
  static mivit_tensor_map_encode_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<mivit_tensor_map_encode_fn>(p);
    else
      (void)cudaGetLastError();
    tried = true;
  }
  if (fn == nullptr) mivit_set_error("cuTensorMapEncodeTiled is not available (no CUDA driver on this machine?)");
  return fn;
}

extern "C" int mivit_abi_version(void) { return MIVIT_ABI_VERSION; }
extern "C" const char* mivit_last_error(void) { return g_err; }
extern "C" int64_t mivit_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" void mivit_reset_launch_count(void) { g_launches.store(0); }
extern "C" void mivit_add_launch_count(int64_t n) { g_launches.fetch_add((long long)n); }

// ---- optional per-kernel device timing (bench.py roofline): CUDA events around tagged launches ----
#include <string.h>
#include <vector>
#include <mutex>

namespace {
struct ProfRec { cudaEvent_t a, b; int tag; double work; };
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfRec> g_prof_recs;
std::vector<cudaEvent_t> g_prof_pool;
char g_prof_names[64][48];
int g_prof_ntags = 0;
}  // namespace

int mivit_prof_tag(const char* name) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < g_prof_ntags; ++i)
    if (strcmp(g_prof_names[i], name) == 0) return i;
  if (g_prof_ntags >= 64) return 63;
  strncpy(g_prof_names[g_prof_ntags], name, 47);
  g_prof_names[g_prof_ntags][47] = 0;
  return g_prof_ntags++;
}
bool mivit_prof_enabled() { return g_prof_on; }
void mivit_prof_begin(int tag, double work, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r;
  r.tag = tag; r.work = work;
  auto get = [&]() {
    cudaEvent_t e;
    if (!g_prof_pool.empty()) { e = g_prof_pool.back(); g_prof_pool.pop_back(); } else { cudaEventCreate(&e); }
    return e;
  };
  r.a = get(); r.b = get();
  cudaEventRecord(r.a, st);
  g_prof_recs.push_back(r);
}
void mivit_prof_end(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof_recs.empty()) cudaEventRecord(g_prof_recs.back().b, st);
}

extern "C" void mivit_profile_enable(int32_t on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
}
// Call after synchronising the stream(s).  Aggregates and clears the records.
extern "C" int32_t mivit_profile_read(mivit_kernel_time* out, int32_t max_entries) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  int n = 0;
  for (int t = 0; t < g_prof_ntags && n < max_entries; ++t) {
    mivit_kernel_time k;
    memset(&k, 0, sizeof(k));
    strncpy(k.name, g_prof_names[t], sizeof(k.name) - 1);
    for (auto& r : g_prof_recs)
      if (r.tag == t) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { k.launches += 1; k.total_ms += ms; k.total_work += r.work; }
      }
    if (k.launches > 0) out[n++] = k;
  }
  for (auto& r : g_prof_recs) { g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b); }
  g_prof_recs.clear();
  return n;
}
