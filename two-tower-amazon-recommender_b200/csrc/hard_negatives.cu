// tfrs.tasks.Retrieval(num_hard_negatives = n)  (layers/loss.py HardNegativeMining): per query row
// the loss runs over the positive and the n highest-scoring negatives only.  The SELECTION is the brute-force top-k
// kernel (k = n + 1, tf.math.top_k order) plus index bookkeeping on the host side; this file evaluates the softmax
// cross-entropy and its gradients on the selected logits, gathered: nq x (n + 1) dot products instead of nq x nc.
// One warp per query row, lanes over the embedding dimension, fp32 arithmetic for both precisions (inputs fp32 or
// bf16).  dC rows are accumulated with fp32 atomics (a candidate is selected by many rows; the order is not fixed,
// like the duplicate rows of the sparse optimizer).
#include "common.cuh"

namespace tt {

template <typename T> __device__ __forceinline__ float ld_elem(const T* p, int64_t i);
template <> __device__ __forceinline__ float ld_elem<float>(const float* p, int64_t i) { return p[i]; }
template <> __device__ __forceinline__ float ld_elem<uint16_t>(const uint16_t* p, int64_t i) { return bf16_bits_to_float(p[i]); }

constexpr int HN_MAX_CHUNKS = 8;      // d <= 256: lane l owns elements l, l + 32, ...

// sel [nq, K] int64: column 0 = the positive, then the selected negatives (-1 = unused slot).
template <typename T>
__global__ void __launch_bounds__(256)
hard_neg_fwd_kernel(const T* __restrict__ q, const T* __restrict__ c, const int64_t* __restrict__ sel, int64_t nq, int K, int d,
                    float inv_temp, const float* __restrict__ w, float* __restrict__ scores, float* __restrict__ row_lse,
                    float* __restrict__ row_pos, float* __restrict__ row_loss) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= nq) return;
  float qv[HN_MAX_CHUNKS];
#pragma unroll
  for (int u = 0; u < HN_MAX_CHUNKS; ++u) qv[u] = (lane + 32 * u < d) ? ld_elem<T>(q, i * d + lane + 32 * u) : 0.f;
  float m = -INFINITY, l = 0.f, pos = 0.f;
  for (int j = 0; j < K; ++j) {
    const int64_t idx = sel[i * K + j];
    float s = -INFINITY;
    if (idx >= 0) {
      float part = 0.f;
#pragma unroll
      for (int u = 0; u < HN_MAX_CHUNKS; ++u)
        if (lane + 32 * u < d) part = fmaf(qv[u], ld_elem<T>(c, idx * d + lane + 32 * u), part);
      s = warp_sum(part) * inv_temp;
      if (s > m) { l = l * expf(m - s) + 1.f; m = s; } else { l += expf(s - m); }
      if (j == 0) pos = s;
    }
    if (lane == 0) scores[i * K + j] = s;
  }
  if (lane == 0) {
    const float lse = m + logf(l);
    row_lse[i] = lse;
    row_pos[i] = pos;
    row_loss[i] = (w ? w[i] : 1.f) * (lse - pos);
  }
}

// ordered sum of the row terms (one block): the SUM loss is bit-reproducible
__global__ void __launch_bounds__(256) hard_neg_loss_sum_kernel(const float* __restrict__ row_loss, int64_t n, float* __restrict__ loss) {
  __shared__ float s_red[256];
  float part = 0.f;
  for (int64_t k = threadIdx.x; k < n; k += 256) part += row_loss[k];
  s_red[threadIdx.x] = part;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = s_red[0];
}

// dq [nq, d] (written), dc [nc, d] (zeroed by the caller, accumulated with atomics)
template <typename T>
__global__ void __launch_bounds__(256)
hard_neg_bwd_kernel(const T* __restrict__ q, const T* __restrict__ c, const int64_t* __restrict__ sel, int64_t nq, int K, int d,
                    float inv_temp, float grad_scale, const float* __restrict__ w, const float* __restrict__ scores,
                    const float* __restrict__ row_lse, float* __restrict__ dq, float* __restrict__ dc) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= nq) return;
  float qv[HN_MAX_CHUNKS], acc[HN_MAX_CHUNKS];
#pragma unroll
  for (int u = 0; u < HN_MAX_CHUNKS; ++u) {
    qv[u] = (lane + 32 * u < d) ? ld_elem<T>(q, i * d + lane + 32 * u) : 0.f;
    acc[u] = 0.f;
  }
  const float lse = row_lse[i];
  const float scale = (w ? w[i] : 1.f) * inv_temp * grad_scale;
  for (int j = 0; j < K; ++j) {
    const int64_t idx = sel[i * K + j];
    if (idx < 0) continue;
    const float g = (expf(scores[i * K + j] - lse) - (j == 0 ? 1.f : 0.f)) * scale;
#pragma unroll
    for (int u = 0; u < HN_MAX_CHUNKS; ++u) {
      if (lane + 32 * u < d) {
        acc[u] = fmaf(g, ld_elem<T>(c, idx * d + lane + 32 * u), acc[u]);
        atomicAdd(dc + idx * d + lane + 32 * u, g * qv[u]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < HN_MAX_CHUNKS; ++u)
    if (lane + 32 * u < d) dq[i * d + lane + 32 * u] = acc[u];
}

static int check_hn(const char* fn, int32_t precision, const void* q, const void* c, const int64_t* sel, int64_t nq, int64_t nc,
                    int64_t K, int64_t d) {
  TT_REQUIRE(precision == TT_F32 || precision == TT_BF16, "%s: unknown precision %d", fn, precision);
  TT_REQUIRE(q && c && sel, "%s: null input", fn);
  TT_REQUIRE(nq >= 0 && nc > 0 && K >= 1 && K < (1 << 30) && d >= 1 && d <= 32 * HN_MAX_CHUNKS, "%s: bad sizes (d <= %d)", fn, 32 * HN_MAX_CHUNKS);
  return TT_OK;
}

}  // namespace tt

using namespace tt;

extern "C" int tt_hard_negative_loss_fwd(int32_t precision, const void* q, const void* c, const int64_t* selected, int64_t nq,
                                         int64_t nc, int64_t k_sel, int64_t d, float inv_temperature, const float* sample_weight,
                                         float* scores, float* row_lse, float* row_pos, float* row_loss, float* loss, void* stream) {
  int rc = check_hn("tt_hard_negative_loss_fwd", precision, q, c, selected, nq, nc, k_sel, d);
  if (rc) return rc;
  TT_REQUIRE(scores && row_lse && row_pos && row_loss && loss, "tt_hard_negative_loss_fwd: null output");
  cudaStream_t st = (cudaStream_t)stream;
  if (nq > 0) {
    const unsigned blocks = (unsigned)ceil_div(nq, 8);
    TT_PROF("hard_neg_fwd_kernel", st);
    if (precision == TT_F32)
      hard_neg_fwd_kernel<float><<<blocks, 256, 0, st>>>((const float*)q, (const float*)c, selected, nq, (int)k_sel, (int)d,
                                                         inv_temperature, sample_weight, scores, row_lse, row_pos, row_loss);
    else
      hard_neg_fwd_kernel<uint16_t><<<blocks, 256, 0, st>>>((const uint16_t*)q, (const uint16_t*)c, selected, nq, (int)k_sel, (int)d,
                                                            inv_temperature, sample_weight, scores, row_lse, row_pos, row_loss);
    TT_LAUNCH_OK("hard_neg_fwd_kernel");
  }
  TT_PROF("hard_neg_loss_sum_kernel", st);
  hard_neg_loss_sum_kernel<<<1, 256, 0, st>>>(row_loss, nq, loss);
  TT_LAUNCH_OK("hard_neg_loss_sum_kernel");
  return TT_OK;
}

extern "C" int tt_hard_negative_loss_bwd(int32_t precision, const void* q, const void* c, const int64_t* selected, int64_t nq,
                                         int64_t nc, int64_t k_sel, int64_t d, float inv_temperature, float grad_scale,
                                         const float* sample_weight, const float* scores, const float* row_lse, float* dq,
                                         float* dc, void* stream) {
  int rc = check_hn("tt_hard_negative_loss_bwd", precision, q, c, selected, nq, nc, k_sel, d);
  if (rc) return rc;
  TT_REQUIRE(scores && row_lse && dq && dc, "tt_hard_negative_loss_bwd: null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  TT_CUDA_OK(cudaMemsetAsync(dc, 0, (size_t)nc * d * sizeof(float), st));
  if (nq == 0) return TT_OK;
  const unsigned blocks = (unsigned)ceil_div(nq, 8);
  TT_PROF("hard_neg_bwd_kernel", st);
  if (precision == TT_F32)
    hard_neg_bwd_kernel<float><<<blocks, 256, 0, st>>>((const float*)q, (const float*)c, selected, nq, (int)k_sel, (int)d, inv_temperature,
                                                       grad_scale, sample_weight, scores, row_lse, dq, dc);
  else
    hard_neg_bwd_kernel<uint16_t><<<blocks, 256, 0, st>>>((const uint16_t*)q, (const uint16_t*)c, selected, nq, (int)k_sel, (int)d,
                                                          inv_temperature, grad_scale, sample_weight, scores, row_lse, dq, dc);
  TT_LAUNCH_OK("hard_neg_bwd_kernel");
  return TT_OK;
}
