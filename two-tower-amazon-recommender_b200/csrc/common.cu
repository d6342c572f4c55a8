// Error TLS, version and device queries behind the C-ABI.
#include "common.cuh"
#include <map>
#include <string>
#include <vector>
#include <algorithm>
#include <string.h>
#include <stdlib.h>

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

}  // namespace tt
// Debug hook: device buffer of 32 int64 = {earliest CTA entry, latest CTA entry/exit} (globaltimer ns) per step
// kernel (ids in common.cuh); fill the even slots with INT64_MAX and the odd ones with 0 before a step.  NULL = off.
extern "C" int tt_debug_timeline(long long* device_buf) {
  tt::set_timeline_retrieval(device_buf);
  tt::set_timeline_tower(device_buf);
  tt::set_timeline_opt(device_buf);
  tt::set_timeline_peer(device_buf);
  return cudaGetLastError() == cudaSuccess ? TT_OK : tt::set_error(TT_ERR_CUDA, "tt_debug_timeline: cudaMemcpyToSymbol failed");
}
namespace tt {

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("TT_NO_PDL");
    on = (e && e[0] && e[0] != '0') ? 0 : 1;
  }
  return on == 1;
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

// ---- per-kernel profiling (eager mode only; never inside graph capture) ---------------
struct ProfRec { const char* name; cudaEvent_t e0, e1; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static cudaStream_t g_prof_stream = nullptr;
static bool g_prof_pending = false;

void prof_begin(const char* name, cudaStream_t st) {
  if (!g_prof_on) return;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
  ProfRec r{name, nullptr, nullptr};
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
  cudaEventRecord(r.e0, st);
  g_prof.push_back(r);
  g_prof_stream = st;
  g_prof_pending = true;
}

void prof_end() {
  if (!g_prof_pending) return;
  cudaEventRecord(g_prof.back().e1, g_prof_stream);
  g_prof_pending = false;
}

}  // namespace tt

extern "C" int tt_profile_enable(int32_t on) {
  for (auto& r : tt::g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  tt::g_prof.clear();
  tt::g_prof_pending = false;
  tt::g_prof_on = on != 0;
  return TT_OK;
}

// Writes "name count total_ms\n" per kernel into host_buf (NUL-terminated); returns bytes needed.
extern "C" int64_t tt_profile_collect(char* host_buf, int64_t buf_len) {
  cudaDeviceSynchronize();
  std::map<std::string, std::pair<long long, double>> agg;
  for (auto& r : tt::g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      auto& a = agg[r.name];
      a.first += 1;
      a.second += ms;
    }
  }
  std::string out;
  char line[256];
  for (auto& kv : agg) {
    snprintf(line, sizeof(line), "%s %lld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  if (host_buf && buf_len > 0) {
    const size_t n = std::min<size_t>(out.size(), (size_t)buf_len - 1);
    memcpy(host_buf, out.data(), n);
    host_buf[n] = 0;
  }
  return (int64_t)out.size() + 1;
}

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
