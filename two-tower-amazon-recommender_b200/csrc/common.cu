// Error TLS, version and device queries behind the C-ABI.
#include "common.cuh"
#include <mutex>

namespace tt {

static thread_local char g_err[1024] = {0};

char* error_buffer() { return g_err; }

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace tt

extern "C" int tt_version(void) { return TT_VERSION; }

extern "C" const char* tt_last_error(void) { return tt::error_buffer(); }

extern "C" int tt_device_check(void) {
  int dev = 0, major = 0;
  TT_CUDA_OK(cudaGetDevice(&dev));
  TT_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10)
    return tt::set_error(TT_ERR_UNSUPPORTED, "libtwotower is built for sm_100a only; device %d has compute capability major %d", dev, major);
  return TT_OK;
}
