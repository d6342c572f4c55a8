// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM as a function of the number of warps (1, 2, 4 per SMSP).
// Decides whether the softmax turn of the retrieval kernels is bounded by pulling S out of tensor memory.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm ldtm.cu && ./ldtm
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>   // 0: ld x32, 1: st x32, 2: ld x32 + 32 dependent FADDs (consume), 3: ld x16
__global__ void __launch_bounds__(512) k(float* out, int iters, long long* cyc, int cols_span) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = lane + i;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t addr = base + (uint32_t)(((warp >> 2) * 32 + (it & 3) * 128) % cols_span);
    if (MODE == 0 || MODE == 2) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
            "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
            "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(addr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[i]);
      }
    } else if (MODE == 1) {
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
          "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
          "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
          ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
            "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
            "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
            "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
          : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(addr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
  }
  const long long t1 = clock64();
  float s = acc;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += __uint_as_float(r[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <int MODE>
void run(const char* name, int warps, int bytes_per_instr) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2048;
  k<MODE><<<148, warps * 32>>>(out, iters, cyc, 512);
  k<MODE><<<148, warps * 32>>>(out, iters, cyc, 512);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  const double bytes = (double)iters * warps * bytes_per_instr;
  printf("%-34s warps/SM %2d: %9.0f cycles, %6.1f cycles per instr per warp, %7.1f B/clk/SM  (%s)\n", name, warps, c, c / iters,
         bytes / c, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16}) run<0>("tcgen05.ld 32x32b.x32 + wait", w, 4096);
  for (int w : {4, 8, 16}) run<3>("tcgen05.ld 32x32b.x16 + wait", w, 2048);
  for (int w : {4, 8, 16}) run<2>("tcgen05.ld x32 + wait + 32 FADD", w, 4096);
  for (int w : {4, 8, 16}) run<1>("tcgen05.st 32x32b.x32 + wait", w, 4096);
  return 0;
}
