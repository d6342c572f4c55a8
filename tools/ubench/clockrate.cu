// Does clock64() tick at the SM clock?  Spin for a fixed number of dependent FMAs, read clock64 and globaltimer.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o clockrate clockrate.cu && ./clockrate
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out, int iters, float* sink) {
  long long g0, g1;
  float x = threadIdx.x;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  long long c0 = clock64();
  for (int i = 0; i < iters; ++i) x = fmaf(x, 1.0001f, 0.5f);      // 4-cycle dependent chain
  long long c1 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
  if (threadIdx.x == 0) { out[0] = c1 - c0; out[1] = g1 - g0; }
  sink[threadIdx.x] = x;
}
int main() {
  long long* out; float* sink; cudaMalloc(&out, 16); cudaMalloc(&sink, 4 * 32);
  for (int rep = 0; rep < 3; ++rep) {
    k<<<1, 32>>>(out, 1 << 20, sink);
    long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("iters=%d  clock64 delta=%lld (%.2f per FMA)  globaltimer delta=%lld ns  => clock64 rate %.3f GHz\n", 1 << 20, h[0],
           (double)h[0] / (1 << 20), h[1], (double)h[0] / h[1]);
  }
  return 0;
}
