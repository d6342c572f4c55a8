// Microbenchmark: exp2 throughput of the softmax inner loop (FFMA -> MUFU.EX2 -> FADD, 128 independent
// elements per thread as in the retrieval kernels) versus the number of warps per SM sub-partition, and
// with a fraction of the exponentials computed by an FMA-pipe polynomial instead of the MUFU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_occ mufu_occ.cu && ./mufu_occ
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 2^x for x <= 0 (softmax arguments): Cody-Waite split with the magic-number round, degree-3 minimax on
// [-0.5, 0.5], exponent patched with an integer add.  Underflows to a denormal/garbage below -126: clamp.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
  float p = fmaf(f, 0.0555041086f, 0.2402265070f);
  p = fmaf(p, f, 0.6931471806f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int POLY_PER_8>
__global__ void __launch_bounds__(512) k(const float* __restrict__ in, float* out, int iters, long long* cyc, float k2) {
  float v[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) v[i] = in[(threadIdx.x * 128 + i) & 4095];
  float l = 0.f, m = 3.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int j = 0; j < 128; j += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float x = fmaf(v[j + u], k2, -m);
        const float e = (u < POLY_PER_8) ? ex2_poly(x) : ex2(x);
        if ((u & 3) == 0) a0 += e; else if ((u & 3) == 1) a1 += e; else if ((u & 3) == 2) a2 += e; else a3 += e;
      }
    }
    l += (a0 + a1) + (a2 + a3);
    m += 1e-7f * l;            // loop-carried so iterations are not hoisted
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int P>
void run(int threads) {
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  cudaMemset(in, 0, 4096 * 4);
  const int iters = 200;
  k<P><<<148, threads>>>(in, out, iters, cyc, 1.3f);
  k<P><<<148, threads>>>(in, out, iters, cyc, 1.3f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  const double per_warp_tile = c / iters;                 // cycles per 128-element row per warp
  printf("poly %d/8  warps/SMSP %d : %7.1f cycles per 128-element row per warp -> %6.1f cycles per row per SMSP (MUFU-only floor 1024)  %s\n",
         P, threads / 128, per_warp_tile, per_warp_tile / (threads / 128), cudaGetErrorString(cudaGetLastError()));
  cudaFree(in); cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int th : {128, 256, 384, 512}) run<0>(th);
  for (int th : {256, 512}) run<1>(th);
  for (int th : {256, 512}) run<2>(th);
  for (int th : {256, 512}) run<3>(th);
  for (int th : {256, 512}) run<4>(th);
  return 0;
}
