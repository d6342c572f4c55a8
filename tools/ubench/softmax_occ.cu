// Microbenchmark: one "tile" of the retrieval forward softmax (128 x 128 logits per SM: scale, running max,
// 2^x, sum) done by 2 warps per SM sub-partition with 128 logits per thread versus 4 warps with 64 logits
// per thread, with every 4th exponential on the FMA pipe.  Decides the warp layout of the softmax stage.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_occ softmax_occ.cu && ./softmax_occ
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
  float p = fmaf(f, 0.05500893f, 0.24221095f);
  p = fmaf(p, f, 0.69328290f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ float fmax3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

template <int N, int POLY_EVERY, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k(const float* __restrict__ in, float* out, int tiles, long long* cyc, float k2) {
  extern __shared__ float4 sm[];                 // THREADS * N floats = 128 KB: every logit is a distinct load
  for (int i = threadIdx.x; i < THREADS * N / 4; i += THREADS) sm[i] = reinterpret_cast<const float4*>(in)[i & 1023];
  float m2 = -1e30f, l = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int t = 0; t < tiles; ++t) {
    float v[N];
#pragma unroll
    for (int i = 0; i < N; i += 4) {               // stands for the TMEM load of the row
      const float4 x = sm[(i / 4) * THREADS + ((threadIdx.x + t) & (THREADS - 1))];   // conflict-free 128-bit reads
      v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
    }
    float cm[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) cm[u] = fmaxf(v[2 * u], v[2 * u + 1]);
#pragma unroll
    for (int j = 8; j < N; j += 8)
#pragma unroll
      for (int u = 0; u < 4; ++u) cm[u] = fmax3(cm[u], v[j + 2 * u], v[j + 2 * u + 1]);
    const float cmax = fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3])) * k2;
    if (cmax > m2) { l *= ex2(m2 - cmax); m2 = cmax; }
    const float nm = -m2;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int j = 0; j < N; j += 4) {
      a0 += ex2(fmaf(v[j], k2, nm));
      a1 += ex2(fmaf(v[j + 1], k2, nm));
      a2 += ex2(fmaf(v[j + 2], k2, nm));
      a3 += (POLY_EVERY == 4) ? ex2_poly(fmaf(v[j + 3], k2, nm)) : ex2(fmaf(v[j + 3], k2, nm));
    }
    l += (a0 + a1) + (a2 + a3);
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l + m2;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int N, int P, int THREADS>
void run(const char* name) {
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 8192 * 4); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  cudaMemset(in, 0, 8192 * 4);
  const int tiles = 256;
  cudaFuncSetAttribute(k<N, P, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, THREADS * N * 4);
  k<N, P, THREADS><<<148, THREADS, THREADS * N * 4>>>(in, out, tiles, cyc, 1.3f);
  k<N, P, THREADS><<<148, THREADS, THREADS * N * 4>>>(in, out, tiles, cyc, 1.3f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  // one kernel "tile" here = THREADS * N logits; normalise to 128 x 128 logits per SM
  printf("%-44s %7.1f cycles per 128x128 logits per SM   (%s)\n", name, c / tiles * (16384.0 / (THREADS * N)), cudaGetErrorString(cudaGetLastError()));
  cudaFree(in); cudaFree(out); cudaFree(cyc);
}

int main() {
  run<128, 0, 256>("2 warps/SMSP x 128 logits, MUFU only");
  run<128, 4, 256>("2 warps/SMSP x 128 logits, 1/4 poly");
  run<64, 0, 512>("4 warps/SMSP x 64 logits, MUFU only");
  run<64, 4, 512>("4 warps/SMSP x 64 logits, 1/4 poly");
  run<32, 4, 1024>("8 warps/SMSP x 32 logits, 1/4 poly");
  return 0;
}
