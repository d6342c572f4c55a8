// K6, second stage -- exact re-rank of the brute-force candidate pool.
//
// The scoring kernels (topk_tc.cu: tcgen05, fp32 accumulation inside the tensor core; topk_simt.cu: an fp32 FMA
// chain) order candidates by a score whose last bits depend on the accumulation order, so on real-valued data
// near-ties may come out in a different order than tf.math.top_k on the reference's score matrix (SURVEY.md A.4,
// section 7 "hard parts").  The order is therefore fixed on an accumulation-order-INDEPENDENT score: the
// scoring stage keeps a pool of k + margin candidates per query, this kernel recomputes the pool's scores
// exactly -- products of bf16 (or fp32) inputs are exact in fp64, the d <= 256 terms are summed in fp64 -- rounds
// them ONCE to fp32 (the dtype of the TFRS score tensor) and sorts by (score desc, candidate index asc).
// A query row is counted in `uncertain` when the margin cannot be shown to have been wide enough: the k-th exact
// score does not clear the pool's lowest stage-one score by 4x the largest stage-one error seen in the pool.
#include "common.cuh"
#include <limits.h>

namespace tt {

template <typename T> __device__ __forceinline__ double rr_ld(const T* p, int64_t i);
template <> __device__ __forceinline__ double rr_ld<float>(const float* p, int64_t i) { return (double)__ldg(p + i); }
template <> __device__ __forceinline__ double rr_ld<uint16_t>(const uint16_t* p, int64_t i) {
  return (double)bf16_bits_to_float(__ldg(p + i));
}

// Candidate-sharded serving: instead of a local [nq, k] result the row of query qi goes straight
// into the receive area of the rank that merges that query -- owner = qi / queries_per_rank -- at list slot [rank]:
// scores at bases[owner] + off_s, ids at bases[owner] + off_i, both laid out [world, queries_per_rank, k].  The stores
// are posted NVLink writes issued by the re-rank epilogue itself (no separate exchange kernel, no staging copy).
struct TopkPeerOut {
  unsigned char* const* bases;     // device array [world] of peer-mapped workspace bases; nullptr = local output
  long long off_s, off_i;
  int queries_per_rank, rank;
};

constexpr int RR_WARPS = 16;     // at most; small query batches use all 16 (latency), large ones 8 warps per CTA
constexpr int RR_MAXJ = 8;      // d <= 256: 8 elements per lane

template <typename T> __device__ __forceinline__ T rr_raw(const T* p, int64_t i) { return __ldg(p + i); }
__device__ __forceinline__ double rr_cvt(float v) { return (double)v; }
__device__ __forceinline__ double rr_cvt(uint16_t v) { return (double)bf16_bits_to_float(v); }

// NJ = elements per lane (d <= 32 * NJ).  One CTA per query: its RR_WARPS warps share the pool, each warp takes four
// entries at a time so that the row loads of four candidates are in flight together (the pool is a random gather
// of 2*d-byte rows); a small query batch still spreads over nq CTAs.
template <typename T, int NJ>
__global__ void __launch_bounds__(RR_WARPS * 32)
topk_rerank_kernel(const T* __restrict__ Q, const T* __restrict__ C, int64_t nq, int d, const float* __restrict__ pool_s,
                   const int64_t* __restrict__ pool_i, int kp, int k, int64_t base, const int64_t* __restrict__ identifiers,
                   float* __restrict__ out_s, int64_t* __restrict__ out_i, int* __restrict__ uncertain, int has_discarded,
                   const TopkPeerOut peer) {
  extern __shared__ __align__(16) uint8_t rr_smem[];
  __shared__ float red_delta[RR_WARPS], red_tmin[RR_WARPS], red_kth;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s32 = reinterpret_cast<float*>(rr_smem);
  int* idx = reinterpret_cast<int*>(s32 + kp);
  const int64_t qi = blockIdx.x;
  double qv[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int e = lane + 32 * j;
    qv[j] = e < d ? rr_ld<T>(Q, qi * d + e) : 0.0;
  }
  if (threadIdx.x == 0) red_kth = -INFINITY;
  float delta = 0.f, tmin = INFINITY;
  const int nwarps = blockDim.x >> 5;
  for (int t0 = 4 * warp; t0 < kp; t0 += 4 * nwarps) {
    int64_t ci[4];
    float approx[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool in = t0 + u < kp;
      ci[u] = in ? __ldg(pool_i + qi * kp + t0 + u) : LLONG_MAX;
      approx[u] = in ? __ldg(pool_s + qi * kp + t0 + u) : 0.f;
    }
    T raw[4][NJ];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int e = lane + 32 * j;
        raw[u][j] = (ci[u] != LLONG_MAX && e < d) ? rr_raw<T>(C, ci[u] * d + e) : T(0);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (t0 + u >= kp) break;
      if (ci[u] == LLONG_MAX) {                  // short pool (fewer candidates than kp): never selected
        if (lane == 0) { s32[t0 + u] = -INFINITY; idx[t0 + u] = INT_MAX; }
        continue;
      }
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc = fma(qv[j], rr_cvt(raw[u][j]), acc);      // fixed order: lane-strided, then the butterfly
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);   // commutative: same value in every lane
      const float f = (float)acc;               // ONE rounding, to nearest even
      delta = fmaxf(delta, fabsf(approx[u] - f));
      tmin = fminf(tmin, approx[u]);
      if (lane == 0) { s32[t0 + u] = f; idx[t0 + u] = (int)ci[u]; }
    }
  }
  if (lane == 0) { red_delta[warp] = delta; red_tmin[warp] = tmin; }
  __syncthreads();
  float* dst_s = out_s + qi * k;
  int64_t* dst_i = out_i + qi * k;
  if (peer.bases != nullptr) {
    const int owner = (int)(qi / peer.queries_per_rank);
    const int64_t slot_row = (int64_t)peer.rank * peer.queries_per_rank + (qi - (int64_t)owner * peer.queries_per_rank);
    dst_s = reinterpret_cast<float*>(peer.bases[owner] + peer.off_s) + slot_row * k;
    dst_i = reinterpret_cast<int64_t*>(peer.bases[owner] + peer.off_i) + slot_row * k;
  }
  for (int e = threadIdx.x; e < kp; e += blockDim.x) {
    const float se = s32[e];
    const int ie = idx[e];
    if (ie == INT_MAX) continue;
    int rank = 0;
    for (int j = 0; j < kp; ++j) {
      const float sj = s32[j];
      const int ij = idx[j];
      rank += (sj > se || (sj == se && ij < ie)) ? 1 : 0;
    }
    if (rank < k) {
      dst_s[rank] = se;
      dst_i[rank] = identifiers ? __ldg(identifiers + ie) : base + ie;
      if (rank == k - 1) red_kth = se;
    }
  }
  if (uncertain != nullptr && has_discarded) {
    __syncthreads();
    if (threadIdx.x == 0) {
      float dl = 0.f, tm = INFINITY;
      for (int w = 0; w < nwarps; ++w) { dl = fmaxf(dl, red_delta[w]); tm = fminf(tm, red_tmin[w]); }
      if (dl > 0.f && !(red_kth > tm + 4.f * dl)) atomicAdd(uncertain, 1);
    }
  }
}

int topk_rerank(int precision, const void* queries, const void* candidates, int64_t nq, int64_t nc, int64_t d,
                const float* pool_s, const int64_t* pool_i, int kp, int k, int64_t base, const int64_t* identifiers,
                float* out_s, int64_t* out_i, int32_t* uncertain, const TopkPeerOut* peer_out, cudaStream_t st) {
  TopkPeerOut peer{};
  if (peer_out) peer = *peer_out;
  TT_REQUIRE(d <= 32 * RR_MAXJ, "tt_topk_bruteforce: exact re-rank needs d <= %d", 32 * RR_MAXJ);
  const size_t smem = (size_t)2 * kp * 4;
  const unsigned blocks = (unsigned)nq;
  const int has_discarded = nc > kp ? 1 : 0;
#define TT_RR_LAUNCH(T, NJ)                                                                                              \
  {                                                                                                                      \
    TT_CUDA_OK(cudaFuncSetAttribute(topk_rerank_kernel<T, NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    TT_PROF("topk_rerank_kernel", st), topk_rerank_kernel<T, NJ><<<blocks, nq >= 2048 ? 256 : RR_WARPS * 32, smem, st>>>(                    \
        (const T*)queries, (const T*)candidates, nq, (int)d, pool_s, pool_i, kp, k, base, identifiers, out_s, out_i,      \
        uncertain, has_discarded, peer);                                                                                 \
  }
  if (precision == TT_F32) {
    if (d <= 128) TT_RR_LAUNCH(float, 4) else TT_RR_LAUNCH(float, 8)
  } else {
    if (d <= 128) TT_RR_LAUNCH(uint16_t, 4) else TT_RR_LAUNCH(uint16_t, 8)
  }
#undef TT_RR_LAUNCH
  TT_LAUNCH_OK("topk_rerank_kernel");
  return TT_OK;
}

}  // namespace tt
