// Library-wide plumbing behind the C ABI: version, thread-local error string, device query.
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace gcs {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return status;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 148;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();   // no device (build container): B200's 148 SMs for size queries
    return 148;
  }
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
    cached_dev = dev;
  }
  return cached;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

struct OpRecord { const char* label; cudaEvent_t a, b; };
static bool g_profile_on = false;
static std::vector<OpRecord> g_records;
static std::mutex g_profile_mu;

ScopedOpTimer::ScopedOpTimer(const char* label, gcs_stream stream) : slot(-1), st(as_stream(stream)) {
  if (!g_profile_on) return;
  std::lock_guard<std::mutex> lk(g_profile_mu);
  OpRecord r{label, nullptr, nullptr};
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, st);
  g_records.push_back(r);
  slot = static_cast<int>(g_records.size()) - 1;
}

ScopedOpTimer::~ScopedOpTimer() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_profile_mu);
  cudaEventRecord(g_records[slot].b, st);
}

}  // namespace gcs

// ---- debug / measurement hooks (not part of the drop-in surface) ----------------------
extern "C" long long gcs_debug_launch_count(void) { return gcs::g_launches.load(); }

extern "C" void gcs_debug_profile_begin(void) {
  std::lock_guard<std::mutex> lk(gcs::g_profile_mu);
  for (auto& r : gcs::g_records) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  gcs::g_records.clear();
  gcs::g_profile_on = true;
}

// Stops profiling, waits for the recorded events and writes "label:count:total_ms;..." into
// out (truncated to cap).  Returns the number of distinct labels.
extern "C" int gcs_debug_profile_end(char* out, int cap) {
  std::lock_guard<std::mutex> lk(gcs::g_profile_mu);
  gcs::g_profile_on = false;
  std::map<std::string, std::pair<int, double>> agg;
  std::vector<std::string> order;
  for (auto& r : gcs::g_records) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      auto it = agg.find(r.label);
      if (it == agg.end()) { agg[r.label] = {1, ms}; order.push_back(r.label); }
      else { it->second.first += 1; it->second.second += ms; }
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  gcs::g_records.clear();
  std::string s;
  for (auto& k : order) {
    char buf[160];
    snprintf(buf, sizeof buf, "%s:%d:%.6f;", k.c_str(), agg[k].first, agg[k].second);
    s += buf;
  }
  if (out && cap > 0) { snprintf(out, cap, "%s", s.c_str()); }
  return static_cast<int>(order.size());
}

extern "C" int gcs_version(void) { return 100; }   // 0.1.0
extern "C" const char* gcs_last_error(void) { return gcs::error_buffer(); }

namespace gcs {
AmaxSink& amax_sink() {
  static thread_local AmaxSink s;
  return s;
}
SyncHook& sync_hook() {
  static thread_local SyncHook h;
  return h;
}
}  // namespace gcs

extern "C" int gcs_set_allreduce_hook(gcs_allreduce_fn fn, void* user, int32_t world_size) {
  if (fn && world_size < 1) return gcs::fail(GCS_ERR_INVALID_ARGUMENT, "gcs_set_allreduce_hook: world_size must be >= 1");
  gcs::SyncHook& h = gcs::sync_hook();
  h.fn = fn;
  h.user = user;
  h.world = fn ? world_size : 1;
  return GCS_OK;
}
extern "C" int gcs_device_sm_count(void) { return gcs::sm_count(); }
