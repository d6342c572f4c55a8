// fp32 CUDA-core tile machinery shared by the fp32-precision retrieval-loss and top-k
// kernels: 64x64 score tiles, 256 threads, 4x4 micro-tiles, operands staged transposed in
// shared memory ([k][64+4]) so the inner loop is two conflict-free LDS.128 + 16 FFMA.
#pragma once
#include "common.cuh"

namespace tt {

constexpr int TS = 64;        // tile side (rows of the stationary operand / streamed operand)
constexpr int TLD = TS + 4;   // leading dimension of the transposed tiles

// Load rows [row0, row0+64) of src[nrows, d] into dst_T[k][row] (and optionally row-major
// dst[row][d+4]).  Rows past nrows are zero-filled.
__device__ __forceinline__ void load_tile(const float* __restrict__ src, int64_t row0, int64_t nrows, int d,
                                          float* __restrict__ dst_T, float* __restrict__ dst) {
  const int c = threadIdx.x & 63;
  const int nq4 = d >> 2;
  const bool ok = row0 + c < nrows;
  const float4* s = reinterpret_cast<const float4*>(src + (row0 + c) * (int64_t)d);
  for (int kq = threadIdx.x >> 6; kq < nq4; kq += 4) {
    float4 v = ok ? __ldg(s + kq) : make_float4(0.f, 0.f, 0.f, 0.f);
    dst_T[(kq * 4 + 0) * TLD + c] = v.x;
    dst_T[(kq * 4 + 1) * TLD + c] = v.y;
    dst_T[(kq * 4 + 2) * TLD + c] = v.z;
    dst_T[(kq * 4 + 3) * TLD + c] = v.w;
    if (dst) *reinterpret_cast<float4*>(dst + c * (d + 4) + kq * 4) = v;
  }
}

// acc[i][j] = sum_k X_T[k][ty*4+i] * Y_T[k][tx*4+j]
__device__ __forceinline__ void tile_dot(const float* __restrict__ X_T, const float* __restrict__ Y_T, int d,
                                         int tx, int ty, float (&acc)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
  for (int k = 0; k < d; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(X_T + k * TLD + ty * 4);
    const float4 b = *reinterpret_cast<const float4*>(Y_T + k * TLD + tx * 4);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

}  // namespace tt
