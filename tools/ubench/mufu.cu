// Microbenchmark: per-SM throughput of ex2 variants (f32, f16x2, bf16x2) and an FMA-pipe
// polynomial exp2, to decide how the softmax warps of the retrieval kernels should compute 2^x.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu mufu.cu && ./mufu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, long long* cyc) {
  float a[8];
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); h[i] = 0xB800B800u + i + threadIdx.x; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 3) {   // Cody-Waite + degree-3 polynomial on the FMA pipe (FA4-style)
        float x = a[i];
        float fl = floorf(x);
        float f = x - fl;
        float p = fmaf(f, 0.05550410866f, 0.2402265070f);
        p = fmaf(p, f, 0.6931471806f);
        p = fmaf(p, f, 1.0f);
        int e = (int)fl;
        a[i] = __int_as_float(__float_as_int(p) + (e << 23)) - 1.5f;
      }
      if (MODE == 4) {   // magic-number variant: no F2I/floor
        float x = a[i];
        float t = x + 12582912.f;              // 1.5 * 2^23: integer part in the low mantissa bits
        float fl = t - 12582912.f;
        float f = x - fl;                       // in [-0.5, 0.5]
        float p = fmaf(f, 0.05550410866f, 0.2402265070f);
        p = fmaf(p, f, 0.6931471806f);
        p = fmaf(p, f, 1.0f);
        a[i] = __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23)) - 1.5f;
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_instr) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  k<MODE><<<148, 512>>>(out, iters, cyc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148, 512>>>(out, iters, cyc);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  double ops = (double)iters * 8 * 512 * per_instr;    // results per SM
  printf("%-28s %8.1f cycles  %6.2f results/clk/SM  (%.3f ms, err=%s)\n", name, c, ops / c, ms, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("poly3 floor/F2I", 1);
  run<4>("poly3 magic", 1);
  return 0;
}
