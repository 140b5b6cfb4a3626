// Library-wide plumbing behind the C ABI: version, thread-local error string, device query.
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

}  // namespace gcs

extern "C" int gcs_version(void) { return 100; }   // 0.1.0
extern "C" const char* gcs_last_error(void) { return gcs::error_buffer(); }
extern "C" int gcs_device_sm_count(void) { return gcs::sm_count(); }
