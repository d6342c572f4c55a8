// Microbenchmark: random 512-byte row reads (one warp per row, one LDG.128 per lane) from tables of
// growing footprint, to separate HBM bandwidth from address-translation effects, and the same rows
// read as [row | slot] pairs from one interleaved allocation vs two separate allocations.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_tlb gather_tlb.cu && ./gather_tlb
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__global__ void gather_rows(const float4* __restrict__ t, const long long* __restrict__ ids, float4* __restrict__ out, int n, int row_f4) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= n) return;
  const long long id = ids[w];
  out[w * 32 + lane] = __ldg(t + id * row_f4 + lane);
}
// two rows per id: either from two tables (a, b) or from one interleaved table (row stride 2x)
__global__ void gather_pairs(const float4* __restrict__ a, const float4* __restrict__ b, const long long* __restrict__ ids,
                             float4* __restrict__ out, int n, int row_f4) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= n) return;
  const long long id = ids[w];
  float4 x = __ldg(a + id * row_f4 + lane), y = __ldg(b + id * row_f4 + lane);
  out[w * 32 + lane] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
}

int main() {
  const int n = 16384;
  float4* out; cudaMalloc(&out, (size_t)n * 512);
  long long* d_ids; cudaMalloc(&d_ids, n * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (double gb : {0.03125, 0.25, 0.5, 1.0, 2.0, 8.0, 32.0}) {
    const long long rows = (long long)(gb * (1ll << 30)) / 512;
    float4* t; if (cudaMalloc(&t, rows * 512) != cudaSuccess) { printf("alloc %.2f GB failed\n", gb); continue; }
    cudaMemset(t, 0, rows * 512);
    float best = 1e9, tot = 0;
    for (int rep = 0; rep < 12; ++rep) {
      std::vector<long long> h(n);
      for (auto& v : h) v = (long long)(((unsigned long long)rand() << 31 | rand()) % (unsigned long long)rows);
      cudaMemcpy(d_ids, h.data(), n * 8, cudaMemcpyHostToDevice);
      cudaEventRecord(e0);
      gather_rows<<<n / 8, 256>>>(t, d_ids, out, n, 32);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep >= 2) { best = ms < best ? ms : best; tot += ms; }
    }
    printf("rows of 512 B: footprint %6.2f GB  n=%d  best %.2f us  avg %.2f us  (%.0f GB/s read at best)\n", gb, n, best * 1e3, tot / 10 * 1e3, n * 512.0 / (best * 1e-3) / 1e9);
    cudaFree(t);
  }
  {
    const long long rows = 1ll << 20;   // 1M rows: 512 MB per array
    float4 *a, *b, *ab;
    cudaMalloc(&a, rows * 512); cudaMalloc(&b, rows * 512); cudaMalloc(&ab, rows * 1024);
    cudaMemset(a, 0, rows * 512); cudaMemset(b, 0, rows * 512); cudaMemset(ab, 0, rows * 1024);
    for (int mode = 0; mode < 2; ++mode) {
      float best = 1e9;
      for (int rep = 0; rep < 12; ++rep) {
        std::vector<long long> h(n);
        for (auto& v : h) v = (long long)(((unsigned long long)rand() << 31 | rand()) % (unsigned long long)rows);
        cudaMemcpy(d_ids, h.data(), n * 8, cudaMemcpyHostToDevice);
        cudaEventRecord(e0);
        if (mode == 0) gather_pairs<<<n / 8, 256>>>(a, b, d_ids, out, n, 32);
        else gather_pairs<<<n / 8, 256>>>(ab, ab + 32, d_ids, out, n, 64);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2) best = ms < best ? ms : best;
      }
      printf("row+slot pairs, 1M ids: %s  best %.2f us\n", mode == 0 ? "two 512 MB arrays     " : "one interleaved 1 GB  ", best * 1e3);
    }
  }
  return 0;
}
