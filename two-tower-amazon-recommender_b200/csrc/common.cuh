// Shared helpers for libtwotower.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/twotower.h"
#include "../../include/twotower_debug.h"

namespace tt {

// thread-local error text behind tt_last_error()
char* error_buffer();
int set_error(int code, const char* fmt, ...);

#define TT_REQUIRE(cond, ...)                                        \
  do {                                                               \
    if (!(cond)) return ::tt::set_error(TT_ERR_INVALID_ARG, __VA_ARGS__); \
  } while (0)

#define TT_CUDA_OK(expr)                                                                   \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return ::tt::set_error(TT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                  \
                             cudaGetErrorString(_e), __FILE__, __LINE__);                  \
  } while (0)

#define TT_LAUNCH_OK(name)                                                                 \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess)                                                                 \
      return ::tt::set_error(TT_ERR_CUDA, "launch of %s failed: %s", name,                 \
                             cudaGetErrorString(_e));                                      \
    ::tt::prof_end();                                                                      \
  } while (0)

// Optional per-kernel timing (tt_profile_enable): a CUDA event pair around a launch on its own
// stream.  TT_PROF(name, stream) goes immediately before the <<<>>>; TT_LAUNCH_OK closes it.
void prof_begin(const char* name, cudaStream_t st);
void prof_end();
#define TT_PROF(name, st) ::tt::prof_begin(name, st)

// Programmatic dependent launch (PDL).  The kernels of one training step form a chain on one stream; launched
// with programmatic stream serialization, kernel N+1 may be scheduled while kernel N drains: its CTAs run their
// prologue (barrier init, TMEM allocation, descriptor prefetch) on SMs that kernel N has left and then block in
// pdl_wait() until kernel N has completed and its writes are visible.  Rules kept by every PDL kernel here:
// nothing before pdl_wait() touches global memory, and EVERY thread executes pdl_wait() (so the completion of
// kernel N+1 implies the completion of kernel N for whoever follows).  TT_NO_PDL=1 falls back to plain launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// In-stream timeline (tt_debug_timeline): kernel k records {earliest CTA entry, latest CTA entry or exit} in
// globaltimer ns at tl[2k], tl[2k+1].  Read through a per-translation-unit __device__ pointer at run time, so it
// also works inside an already captured CUDA graph.  Kernel ids: 0 tower fwd, 1 loss fwd, 2 dQ, 3 dC,
// 4 tower bwd, 5 optimizer step, 6 sparse prepare, 7 combine partials, 8 fold dense parts, 9-14 peer.cu.
#define TT_TL_DEFINE(setter)                                                              \
  static __device__ long long* g_tl = nullptr;                                            \
  void setter(long long* p) { cudaMemcpyToSymbol(g_tl, &p, sizeof(p)); }
__device__ __forceinline__ void tl_mark(long long* tl, int k, bool entry) {
  if (tl && threadIdx.x == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (entry) atomicMin(&tl[2 * k], t);
    atomicMax(&tl[2 * k + 1], t);
  }
}
void set_timeline_retrieval(long long* p);
void set_timeline_tower(long long* p);
void set_timeline_opt(long long* p);
void set_timeline_peer(long long* p);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

int num_sms();

__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) {
  return __uint_as_float(static_cast<uint32_t>(b) << 16);
}
__device__ __forceinline__ uint16_t float_to_bf16_bits(float f) {
  return __bfloat16_as_ushort(__float2bfloat16_rn(f));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// tfrs/layers/loss.py: MIN_FLOAT = np.finfo(np.float32).min / 100.0
#define TT_MIN_FLOAT (-3.4028234663852886e36f)

}  // namespace tt
