// K2 (fp32 precision) -- tf.keras.layers.Dense forward/backward on CUDA-core FFMA.
// This is the 1e-5-parity path: tcgen05 has no true-fp32 MMA, so precision TT_F32 runs a
// classic 64x64x16 shared-memory tiled SGEMM with generic operand strides (NN for forward,
// NT for dgrad, TN for wgrad) and fused bias / ReLU / ReLU-mask epilogues.  The bf16
// tensor-core path (precision TT_BF16) lives in dense_tc.cu.
#include "common.cuh"

namespace tt {

int tc_dense_fwd(const void* x, const void* kernel, const float* bias, void* y, float* y_f32,
                 int64_t M, int64_t in_dim, int64_t out_dim, int relu, cudaStream_t stream);
int tc_dense_bwd(const void* dy, const void* x, const void* kernel, void* dx, float* dx_f32,
                 float* dkernel_parts, int num_parts, float* dbias_parts, int64_t M, int64_t in_dim,
                 int64_t out_dim, int relu_mask_x, cudaStream_t stream);
int tc_dense_bwd_num_parts(int64_t M, int64_t in_dim, int64_t out_dim);

constexpr int BM = 64, BN = 64, BK = 16;

// C[M,N] = epi( sum_k A(m,k) * B(k,n) ), A(m,k) = A[m*sa_m + k*sa_k], B(k,n) = B[k*sb_k + n*sb_n]
// epi: + bias[n]; relu; * (mask[m*ldc+n] > 0)
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, int64_t sa_m, int64_t sa_k, const float* __restrict__ B,
             int64_t sb_k, int64_t sb_n, float* __restrict__ C, int64_t ldc, int M, int N, int K,
             const float* __restrict__ bias, int relu, const float* __restrict__ mask) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + i * 256;
      int m, k;
      if (sa_k == 1) { m = idx >> 4; k = idx & 15; } else { m = idx & 63; k = idx >> 6; }
      float v = 0.f;
      if (m0 + m < M && k0 + k < K) v = A[(int64_t)(m0 + m) * sa_m + (int64_t)(k0 + k) * sa_k];
      As[k][m] = v;
      int n, kk;
      if (sb_n == 1) { n = idx & 63; kk = idx >> 6; } else { kk = idx & 15; n = idx >> 4; }
      v = 0.f;
      if (n0 + n < N && k0 + kk < K) v = B[(int64_t)(k0 + kk) * sb_k + (int64_t)(n0 + n) * sb_n];
      Bs[kk][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      if (relu) v = fmaxf(v, 0.f);
      if (mask && !(mask[(int64_t)m * ldc + n] > 0.f)) v = 0.f;
      C[(int64_t)m * ldc + n] = v;
    }
  }
}

// out[p][n] = sum over the p-th row slice of X[m, n], fixed order (deterministic).
// grid (ceil(N/32), num_parts); 256 threads = 32 columns x 8 row groups.
__global__ void __launch_bounds__(256)
colsum_f32_kernel(const float* __restrict__ X, float* __restrict__ out, int64_t M, int64_t N, int64_t rows_per_part) {
  __shared__ float part[8][32];
  const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int64_t n = (int64_t)blockIdx.x * 32 + c;
  const int64_t m_lo = (int64_t)blockIdx.y * rows_per_part, m_hi = min(M, m_lo + rows_per_part);
  float s = 0.f;
  if (n < N)
    for (int64_t m = m_lo + g; m < m_hi; m += 8) s += X[m * N + n];
  part[g][c] = s;
  __syncthreads();
  if (g == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][c];
    out[(int64_t)blockIdx.y * N + n] = t;
  }
}

// out[i] = sum_p parts[p][i]
__global__ void __launch_bounds__(256)
sum_parts_kernel(const float* __restrict__ parts, int num_parts, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int p = 0; p < num_parts; ++p) s += parts[(int64_t)p * n + i];
  out[i] = s;
}

__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ in, uint16_t* __restrict__ out, int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(in + i);
    *reinterpret_cast<uint2*>(out + i) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  } else {
    for (int64_t k = i; k < n; ++k) out[k] = float_to_bf16_bits(in[k]);
  }
}

static int sgemm(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k, int64_t sb_n,
                 float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, const float* bias, int relu,
                 const float* mask, cudaStream_t stream) {
  dim3 grid((unsigned)ceil_div(N, BN), (unsigned)ceil_div(M, BM));
  TT_PROF("sgemm_kernel", stream);
  sgemm_kernel<<<grid, 256, 0, stream>>>(A, sa_m, sa_k, B, sb_k, sb_n, C, ldc, (int)M, (int)N, (int)K, bias, relu, mask);
  TT_LAUNCH_OK("sgemm_kernel");
  return TT_OK;
}

}  // namespace tt

using namespace tt;

extern "C" int tt_dense_fwd(int32_t precision, const void* x, const void* kernel, const float* bias,
                            void* y, float* y_f32, int64_t M, int64_t in_dim, int64_t out_dim,
                            int32_t relu, void* stream) {
  TT_REQUIRE(x && kernel && y, "tt_dense_fwd: null buffer");
  TT_REQUIRE(M > 0 && in_dim > 0 && out_dim > 0 && M < (1ll << 31), "tt_dense_fwd: bad sizes");
  if (precision == TT_F32) {
    TT_REQUIRE(y_f32 == nullptr, "tt_dense_fwd: y_f32 is a bf16-path output");
    return sgemm((const float*)x, in_dim, 1, (const float*)kernel, out_dim, 1, (float*)y, out_dim, M, out_dim,
                 in_dim, bias, relu, nullptr, (cudaStream_t)stream);
  }
  TT_REQUIRE(precision == TT_BF16, "tt_dense_fwd: unknown precision %d", precision);
  return tc_dense_fwd(x, kernel, bias, y, y_f32, M, in_dim, out_dim, relu, (cudaStream_t)stream);
}

extern "C" int32_t tt_dense_bwd_num_parts(int32_t precision, int64_t M, int64_t in_dim, int64_t out_dim) {
  if (precision == TT_F32) return 1;
  return tc_dense_bwd_num_parts(M, in_dim, out_dim);
}

extern "C" int tt_dense_bwd(int32_t precision, const void* dy, const void* x, const void* kernel, void* dx,
                            float* dx_f32, float* dkernel_parts, int32_t num_parts, float* dbias_parts,
                            int64_t M, int64_t in_dim, int64_t out_dim, int32_t relu_mask_x, void* stream) {
  TT_REQUIRE(dy && x && kernel && dkernel_parts, "tt_dense_bwd: null buffer");
  TT_REQUIRE(M > 0 && in_dim > 0 && out_dim > 0 && M < (1ll << 31), "tt_dense_bwd: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == TT_F32) {
    TT_REQUIRE(num_parts == 1, "tt_dense_bwd: fp32 path writes a single gradient part");
    TT_REQUIRE(dx_f32 == nullptr, "tt_dense_bwd: dx_f32 is a bf16-path output");
    int rc;
    if (dx) {   // dx = dy[M,out] @ kernel[in,out]^T  (* relu mask of x)
      rc = sgemm((const float*)dy, out_dim, 1, (const float*)kernel, 1, out_dim, (float*)dx, in_dim, M, in_dim,
                 out_dim, nullptr, 0, relu_mask_x ? (const float*)x : nullptr, st);
      if (rc) return rc;
    }
    // dkernel[in,out] = x[M,in]^T @ dy[M,out]
    rc = sgemm((const float*)x, 1, in_dim, (const float*)dy, out_dim, 1, dkernel_parts, out_dim, in_dim, out_dim,
               M, nullptr, 0, nullptr, st);
    if (rc) return rc;
    if (dbias_parts) {
      TT_PROF("colsum_f32_kernel", st);
      colsum_f32_kernel<<<(unsigned)ceil_div(out_dim, 32), 256, 0, st>>>((const float*)dy, dbias_parts, M, out_dim, M);
      TT_LAUNCH_OK("colsum_f32_kernel");
    }
    return TT_OK;
  }
  TT_REQUIRE(precision == TT_BF16, "tt_dense_bwd: unknown precision %d", precision);
  return tc_dense_bwd(dy, x, kernel, dx, dx_f32, dkernel_parts, num_parts, dbias_parts, M, in_dim, out_dim,
                      relu_mask_x, st);
}

extern "C" int tt_colsum_f32(const float* x, float* out_parts, int64_t rows, int64_t cols, int32_t num_parts,
                             void* stream) {
  TT_REQUIRE(x && out_parts && rows > 0 && cols > 0 && num_parts >= 1, "tt_colsum_f32: bad arguments");
  const int64_t rpp = ceil_div(rows, num_parts);
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)num_parts);
  TT_PROF("colsum_f32_kernel", (cudaStream_t)stream);
  colsum_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, out_parts, rows, cols, rpp);
  TT_LAUNCH_OK("colsum_f32_kernel");
  return TT_OK;
}

extern "C" int tt_sum_parts_f32(const float* parts, int32_t num_parts, int64_t n, float* out, void* stream) {
  TT_REQUIRE(parts && out && num_parts >= 1 && n > 0, "tt_sum_parts_f32: bad arguments");
  TT_PROF("sum_parts_kernel", (cudaStream_t)stream);
  sum_parts_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(parts, num_parts, n, out);
  TT_LAUNCH_OK("sum_parts_kernel");
  return TT_OK;
}

extern "C" int tt_cast_f32_to_bf16(const float* in, uint16_t* out, int64_t n, void* stream) {
  TT_REQUIRE(in && out && n > 0, "tt_cast_f32_to_bf16: bad arguments");
  TT_REQUIRE(aligned16(in) && (reinterpret_cast<uintptr_t>(out) & 7u) == 0, "tt_cast_f32_to_bf16: unaligned buffers");
  TT_PROF("cast_bf16_kernel", (cudaStream_t)stream);
  cast_bf16_kernel<<<(unsigned)ceil_div(ceil_div(n, 4), 256), 256, 0, (cudaStream_t)stream>>>(in, out, n);
  TT_LAUNCH_OK("cast_bf16_kernel");
  return TT_OK;
}
