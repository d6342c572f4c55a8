// Microbenchmark: the per-tile softmax inner loops of the retrieval kernels with packed fp32x2 math
// (FFMA2 / FADD2, sm_100) against the scalar form, at several MUFU : FMA-pipe-polynomial splits.
// FWD = scale, running max, 2^x, row sum.  BWD = scale - lse, 2^x, pack to bf16x2.
// 2 warps per SM sub-partition x 128 logits per thread, like the kernels.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_pk softmax_pk.cu && ./softmax_pk
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
__device__ __forceinline__ uint64_t pk2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }

// 2^x for a packed pair on the FMA pipe
__device__ __forceinline__ uint64_t ex2_poly2(uint64_t x2) {
  float x0, x1; up2(x2, x0, x1);
  x0 = fmaxf(x0, -126.f); x1 = fmaxf(x1, -126.f);
  x2 = pk2(x0, x1);
  const uint64_t t = add2(x2, pk2(12582912.f, 12582912.f));
  const uint64_t f = sub2(x2, add2(t, pk2(-12582912.f, -12582912.f)));
  uint64_t p = fma2(f, pk2(0.05500893f, 0.05500893f), pk2(0.24221095f, 0.24221095f));
  p = fma2(p, f, pk2(0.69328290f, 0.69328290f));
  p = fma2(p, f, pk2(1.f, 1.f));
  float p0, p1, t0, t1; up2(p, p0, p1); up2(t, t0, t1);
  return pk2(__int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23)),
             __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23)));
}
__device__ __forceinline__ uint64_t ex2_mufu2(uint64_t x2) {
  float x0, x1; up2(x2, x0, x1);
  return pk2(ex2(x0), ex2(x1));
}

// MODE 0: forward scalar, every 4th on poly (the r01 kernel).  MODE 1: forward packed, PP of every 4 pairs on poly.
// MODE 2: backward scalar 1/4 poly.  MODE 3: backward packed.
template <int MODE, int PP>
__global__ void __launch_bounds__(256, 1) k(const float* __restrict__ in, float* out, int tiles, long long* cyc, float k2) {
  constexpr int N = 128, THREADS = 256;
  extern __shared__ float4 sm[];
  for (int i = threadIdx.x; i < THREADS * N / 4; i += THREADS) sm[i] = reinterpret_cast<const float4*>(in)[i & 1023];
  float m2 = -1e30f, l = 0.f;
  uint32_t sink = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int t = 0; t < tiles; ++t) {
    float v[N];
#pragma unroll
    for (int i = 0; i < N; i += 4) {
      const float4 x = sm[(i / 4) * THREADS + ((threadIdx.x + t) & (THREADS - 1))];
      v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
    }
    if (MODE <= 1) {
      float cm[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) cm[u] = fmaxf(v[2 * u], v[2 * u + 1]);
#pragma unroll
      for (int j = 8; j < N; j += 8)
#pragma unroll
        for (int u = 0; u < 4; ++u) cm[u] = fmax3(cm[u], v[j + 2 * u], v[j + 2 * u + 1]);
      const float cmax = fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3])) * k2;
      if (cmax > m2) { l *= ex2(m2 - cmax); m2 = cmax; }
    }
    const float nm = MODE <= 1 ? -m2 : -3.f;
    if (MODE == 0) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int j = 0; j < N; j += 4) {
        a0 += ex2(fmaf(v[j], k2, nm)); a1 += ex2(fmaf(v[j + 1], k2, nm));
        a2 += ex2(fmaf(v[j + 2], k2, nm)); a3 += ex2_poly(fmaf(v[j + 3], k2, nm));
      }
      l += (a0 + a1) + (a2 + a3);
    } else if (MODE == 1) {
      const uint64_t K2 = pk2(k2, k2), NM = pk2(nm, nm);
      uint64_t acc0 = pk2(0.f, 0.f), acc1 = pk2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < N; j += 8) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint64_t x2 = fma2(pk2(v[j + 2 * u], v[j + 2 * u + 1]), K2, NM);
          const uint64_t e2 = (u < PP) ? ex2_poly2(x2) : ex2_mufu2(x2);
          if (u & 1) acc1 = add2(acc1, e2); else acc0 = add2(acc0, e2);
        }
      }
      float s0, s1, s2, s3; up2(acc0, s0, s1); up2(acc1, s2, s3);
      l += (s0 + s1) + (s2 + s3);
    } else if (MODE == 2) {
#pragma unroll
      for (int j = 0; j < N; j += 4) {
        const float p0 = ex2(fmaf(v[j], k2, nm)), p1 = ex2(fmaf(v[j + 1], k2, nm));
        const float p2 = ex2(fmaf(v[j + 2], k2, nm)), p3 = ex2_poly(fmaf(v[j + 3], k2, nm));
        sink ^= pack_bf16x2(p0, p1) + pack_bf16x2(p2, p3);
      }
    } else {
      const uint64_t K2 = pk2(k2, k2), NM = pk2(nm, nm);
#pragma unroll
      for (int j = 0; j < N; j += 8) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint64_t x2 = fma2(pk2(v[j + 2 * u], v[j + 2 * u + 1]), K2, NM);
          const uint64_t e2 = (u < PP) ? ex2_poly2(x2) : ex2_mufu2(x2);
          float e0, e1; up2(e2, e0, e1);
          sink ^= pack_bf16x2(e0, e1);
        }
      }
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l + m2 + __uint_as_float(sink);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE, int PP>
void run(const char* name) {
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 8192 * 4); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  cudaMemset(in, 0, 8192 * 4);
  const int tiles = 256;
  cudaFuncSetAttribute(k<MODE, PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 128 * 4);
  k<MODE, PP><<<148, 256, 256 * 128 * 4>>>(in, out, tiles, cyc, 1.3f);
  k<MODE, PP><<<148, 256, 256 * 128 * 4>>>(in, out, tiles, cyc, 1.3f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  // one iteration = 256 threads x 128 logits = two 128x128 tiles
  printf("%-40s %7.1f cycles per 128x128 logits per SM   (%s)\n", name, c / tiles / 2, cudaGetErrorString(cudaGetLastError()));
  cudaFree(in); cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0, 0>("fwd scalar, 1/4 poly (r01)");
  run<1, 0>("fwd packed, MUFU only");
  run<1, 1>("fwd packed, 1/4 poly");
  run<1, 2>("fwd packed, 2/4 poly");
  run<1, 3>("fwd packed, 3/4 poly");
  run<2, 0>("bwd scalar, 1/4 poly (r01)");
  run<3, 0>("bwd packed, MUFU only");
  run<3, 1>("bwd packed, 1/4 poly");
  run<3, 2>("bwd packed, 2/4 poly");
  run<3, 3>("bwd packed, 3/4 poly");
  return 0;
}
