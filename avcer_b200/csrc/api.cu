// Error reporting, version and device checks of the C ABI.
#include "common.h"

#include <string.h>

namespace avcer {

static thread_local char g_err[1024] = "";

int set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

int& sm_limit_ref() {
  static thread_local int limit = 0;
  return limit;
}

}  // namespace avcer

extern "C" int avcer_set_sm_limit(int n_sms) {
  if (n_sms < 0 || (n_sms & 1)) return avcer::set_error("set_sm_limit: %d must be 0 (all) or a positive even SM count", n_sms);
  avcer::sm_limit_ref() = n_sms;
  return 0;
}

extern "C" const char* avcer_last_error(void) { return avcer::g_err; }

extern "C" int avcer_version(void) { return 100; }

extern "C" const char* avcer_storage_type(void) { return AVCER_STORAGE_NAME; }

extern "C" int avcer_device_check(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return avcer::set_error("no CUDA device visible (%s); avcer_b200 has no CPU fallback",
                            e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  int dev = 0, major = 0;
  AVCER_CUDA(cudaGetDevice(&dev));
  AVCER_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10)
    return avcer::set_error("device compute capability %d.x is not sm_100a (B200); kernels are sm_100a only", major);
  return 0;
}

extern "C" int avcer_num_sms(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return n;
}
