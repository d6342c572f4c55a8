// K6 (fp32 precision) -- brute-force scoring fused with a running top-k; list merge; hit@k.
// Restates tfrs.layers.factorized_top_k.BruteForce.call (matmul + tf.math.top_k: sorted
// descending, ties -> lower index first), faiss.IndexFlatIP.search and
// tfrs.metrics.FactorizedTopK.  The [nq, nc] score matrix is never
// written: each CTA keeps 64 queries stationary, streams a candidate range in 64-row tiles,
// filters scores against the row's current k-th best and folds survivors into a sorted
// shared-memory list (topk_select.cuh).  bf16/tcgen05 scoring is topk_tc.cu.
#include "simt_tile.cuh"
#include "topk_select.cuh"
#include <limits.h>

namespace tt {

int tc_topk(const void* queries, const void* candidates, int64_t nq, int64_t nc, int64_t d, int k,
            int64_t cand_index_base, const int64_t* identifiers, float* out_scores, int64_t* out_ids, void* ws,
            int64_t ws_bytes, cudaStream_t st);
int tc_topk_num_splits(int64_t nq, int64_t nc, int64_t d, int k);
int tc_topk_max_k(int64_t d);
int64_t tc_topk_stage_bytes(int64_t nq, int64_t nc, int64_t d, int kp);
int tc_topk_scan_launches(int64_t nq, int64_t nc, int64_t d, int kp);

// Partial (or final) result writer shared with the tensor-core kernel: row r of the state ->
// out arrays, padding short lists with (-inf, INT64_MAX).
__device__ __forceinline__ void topk_write_row(const TopkRowState& st, int r, int lane, float* out_s,
                                               int64_t* out_i, int64_t base, const int64_t* identifiers) {
  const int k = st.k, cnt = st.count[r];
  for (int t = lane; t < k; t += 32) {
    if (t < cnt) {
      const int idx = st.list_i[(size_t)r * k + t];
      out_s[t] = st.list_s[(size_t)r * k + t];
      out_i[t] = identifiers ? __ldg(identifiers + idx) : base + idx;
    } else {
      out_s[t] = -INFINITY;
      out_i[t] = LLONG_MAX;
    }
  }
}

template <int KU>
__global__ void __launch_bounds__(256)
topk_simt_kernel(const float* __restrict__ Q, const float* __restrict__ C, int64_t nq, int64_t nc, int d, int k,
                 int64_t cand_base, const int64_t* __restrict__ identifiers, int64_t split_len,
                 float* __restrict__ out_s, int64_t* __restrict__ out_i) {
  extern __shared__ __align__(16) float smem[];
  float* Xs_T = smem;
  float* Ys_T = Xs_T + d * TLD;
  TopkRowState st = topk_state_carve(Ys_T + d * TLD, TS, k, TS);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, warp = tid >> 5;
  const int64_t q0 = (int64_t)blockIdx.x * TS;
  const int64_t c_lo = (int64_t)blockIdx.y * split_len;
  const int64_t c_hi = min(nc, c_lo + split_len);

  if (tid < TS) { st.count[tid] = 0; st.tau[tid] = -INFINITY; st.pend_n[tid] = 0; }
  load_tile(Q, q0, nq, d, Xs_T, nullptr);

  for (int64_t y0 = c_lo; y0 < c_hi; y0 += TS) {
    __syncthreads();
    load_tile(C, y0, c_hi, d, Ys_T, nullptr);
    __syncthreads();
    float acc[4][4];
    tile_dot(Xs_T, Ys_T, d, tx, ty, acc);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty * 4 + i;
      if (q0 + r >= nq) continue;
      const float tau = st.tau[r];
      const bool full = st.count[r] == k;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t ci = y0 + tx * 4 + j;
        if (ci >= c_hi) continue;
        const float s = acc[i][j];
        if (!full || s > tau) {
          const int slot = atomicAdd(&st.pend_n[r], 1);
          st.pend_s[r * TS + slot] = s;
          st.pend_i[r * TS + slot] = (int)ci;
        }
      }
    }
    __syncthreads();
    for (int r = warp * 8; r < warp * 8 + 8; ++r) topk_merge_row<KU>(st, r, lane);
  }
  __syncthreads();
  for (int r = warp * 8; r < warp * 8 + 8; ++r) {
    if (q0 + r >= nq) continue;
    const size_t o = ((size_t)blockIdx.y * nq + (q0 + r)) * k;
    topk_write_row(st, r, lane, out_s + o, out_i + o, gridDim.y > 1 ? 0 : cand_base, gridDim.y > 1 ? nullptr : identifiers);
  }
}

// L-way merge of sorted lists, one warp per query; lane l walks lists l, l + 32, ... (L <= 32 * MERGE_LPL).
constexpr int MERGE_LPL = 8;
__global__ void __launch_bounds__(256)
topk_merge_kernel(const float* __restrict__ scores, const int64_t* __restrict__ ids, int L, int64_t nq, int k_in,
                  int k_out, int64_t base, const int64_t* __restrict__ identifiers, float* __restrict__ out_s,
                  int64_t* __restrict__ out_i, const int* __restrict__ run_if) {
  if (run_if != nullptr && *run_if == 0) return;
  const int lane = threadIdx.x & 31;
  const int64_t qi = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (qi >= nq) return;
  int head[MERGE_LPL];
  float hs[MERGE_LPL];
  int64_t hi[MERGE_LPL];
#pragma unroll
  for (int u = 0; u < MERGE_LPL; ++u) {
    const int list = lane + 32 * u;
    head[u] = 0;
    hs[u] = -INFINITY; hi[u] = LLONG_MAX;
    if (list < L && k_in > 0) {
      hs[u] = scores[((size_t)list * nq + qi) * k_in];
      hi[u] = ids[((size_t)list * nq + qi) * k_in];
    }
  }
  for (int t = 0; t < k_out; ++t) {
    // this lane's best head (ties: lower index, then lower list), then the warp's
    float bs = hs[0]; int64_t bi = hi[0]; int bl = lane;
#pragma unroll
    for (int u = 1; u < MERGE_LPL; ++u) {
      const bool take = hs[u] > bs || (hs[u] == bs && hi[u] < bi);
      if (take) { bs = hs[u]; bi = hi[u]; bl = lane + 32 * u; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float s2 = __shfl_xor_sync(0xffffffffu, bs, o);
      const int64_t i2 = __shfl_xor_sync(0xffffffffu, bi, o);
      const int l2 = __shfl_xor_sync(0xffffffffu, bl, o);
      const bool take = s2 > bs || (s2 == bs && (i2 < bi || (i2 == bi && l2 < bl)));
      if (take) { bs = s2; bi = i2; bl = l2; }
    }
    if (lane == 0) {
      out_s[qi * k_out + t] = bs;
      out_i[qi * k_out + t] = (bi == LLONG_MAX) ? bi : (identifiers ? __ldg(identifiers + bi) : base + bi);
    }
    if ((bl & 31) == lane) {
#pragma unroll
      for (int u = 0; u < MERGE_LPL; ++u) {
        if (bl == lane + 32 * u) {
          ++head[u];
          if (head[u] < k_in) {
            hs[u] = scores[((size_t)bl * nq + qi) * k_in + head[u]];
            hi[u] = ids[((size_t)bl * nq + qi) * k_in + head[u]];
          } else { hs[u] = -INFINITY; hi[u] = LLONG_MAX; }
        }
      }
    }
  }
}

// FactorizedTopK hit counting (score mode / id mode).
__global__ void __launch_bounds__(256)
topk_hits_kernel(const float* __restrict__ positive, const float* __restrict__ topk_scores,
                 const int64_t* __restrict__ topk_ids, const int64_t* __restrict__ true_ids,
                 const float* __restrict__ w, int64_t nq, int k, const int* __restrict__ ks_dev_unused,
                 int ks0, int ks1, int ks2, int ks3, int ks4, int ks5, int ks6, int ks7, int num_ks,
                 float* __restrict__ hits_out, float* __restrict__ weight_out) {
  const int ks[8] = {ks0, ks1, ks2, ks3, ks4, ks5, ks6, ks7};
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float hit[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) hit[j] = 0.f;
  float wi = 0.f;
  if (i < nq) {
    wi = w ? w[i] : 1.f;
    if (true_ids == nullptr) {
      const float p = positive[i];
      int greater = 0;
      for (int t = 0; t < k; ++t) greater += topk_scores[i * k + t] > p ? 1 : 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) if (j < num_ks && greater < ks[j]) hit[j] = wi;
    } else {
      const int64_t tid_ = true_ids[i];
      int first = INT_MAX;
      for (int t = 0; t < k; ++t) if (topk_ids[i * k + t] == tid_) { first = t; break; }
#pragma unroll
      for (int j = 0; j < 8; ++j) if (j < num_ks && first < ks[j]) hit[j] = wi;
    }
  }
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (j >= num_ks) break;
    float v = warp_sum(hit[j]);
    if (lane == 0 && v != 0.f) atomicAdd(hits_out + j, v);
  }
  float ws = warp_sum(wi);
  if (lane == 0 && ws != 0.f) atomicAdd(weight_out, ws);
}

template <typename T> __device__ __forceinline__ float ld_as_float(const T* p, int64_t i);
template <> __device__ __forceinline__ float ld_as_float<float>(const float* p, int64_t i) { return p[i]; }
template <> __device__ __forceinline__ float ld_as_float<uint16_t>(const uint16_t* p, int64_t i) { return bf16_bits_to_float(p[i]); }

template <typename T>
__global__ void __launch_bounds__(256)
rowwise_dot_kernel(const T* __restrict__ q, const T* __restrict__ c, float* __restrict__ out, int64_t n, int64_t d) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  float s = 0.f;
  for (int64_t k = lane; k < d; k += 32) s = fmaf(ld_as_float(q, r * d + k), ld_as_float(c, r * d + k), s);
  s = warp_sum(s);
  if (lane == 0) out[r] = s;
}

static size_t simt_topk_smem(int d, int k) { return (size_t)2 * d * TLD * 4 + topk_state_bytes(TS, k, TS); }

}  // namespace tt

namespace tt {
int topk_merge_launch(const float* scores, const int64_t* ids, int num_lists, int64_t nq, int k_in, int k_out,
                      int64_t index_base, const int64_t* identifiers, float* out_scores, int64_t* out_ids,
                      const int* run_if, cudaStream_t st) {
  TT_PROF("topk_merge_kernel", st);
  topk_merge_kernel<<<(unsigned)ceil_div(nq, 8), 256, 0, st>>>(scores, ids, num_lists, nq, k_in, k_out, index_base,
                                                               identifiers, out_scores, out_ids, run_if);
  TT_LAUNCH_OK("topk_merge_kernel");
  return TT_OK;
}
}  // namespace tt

using namespace tt;

static const size_t kMaxSmem = 227 * 1024;

static int simt_num_splits(int64_t nq, int64_t nc) {
  const int64_t qtiles = ceil_div(nq, TS);
  const int64_t target = 2 * (int64_t)num_sms();
  int64_t s = qtiles >= target ? 1 : ceil_div(target, qtiles);
  const int64_t max_by_len = std::max<int64_t>(1, nc / 4096);   // keep >= 4096 candidates per split
  if (s > max_by_len) s = max_by_len;
  if (s > 32) s = 32;
  return (int)s;
}

extern "C" int32_t tt_topk_num_splits(int32_t precision, int64_t nq, int64_t nc, int64_t d, int32_t k) {
  if (precision == TT_BF16) return tc_topk_num_splits(nq, nc, d, k);
  return simt_num_splits(nq, nc);
}

// Pool size of the scoring stage: k + margin candidates per query go to the exact re-rank (topk_rerank.cu).
static int pool_k(int32_t precision, int64_t d, int k, int64_t nc) {
  int64_t kp = std::min<int64_t>(k + TT_TOPK_RERANK_MARGIN, 512);
  if (precision == TT_BF16) kp = std::min<int64_t>(kp, tc_topk_max_k(d));     // the margin shrinks before k is refused
  if (kp < k) kp = k;
  return (int)std::min<int64_t>(kp, nc);
}

static int64_t stage1_bytes(int32_t precision, int64_t nq, int64_t nc, int64_t d, int kp) {
  if (precision == TT_BF16) return tc_topk_stage_bytes(nq, nc, d, kp);
  const int s = tt_topk_num_splits(precision, nq, nc, d, kp);
  if (s <= 1) return 256;
  return round_up((int64_t)s * nq * kp * 4, 256) + round_up((int64_t)s * nq * kp * 8, 256);
}

extern "C" int32_t tt_topk_num_launches(int32_t precision, int64_t nq, int64_t nc, int64_t d, int32_t k) {
  const int kp = pool_k(precision, d, k, nc);
  int n = (tt_topk_num_splits(precision, nq, nc, d, kp) > 1 ? 2 : 1) + 1;      // scoring (+ merge) + re-rank
  if (precision == TT_BF16) n += tc_topk_scan_launches(nq, nc, d, kp);         // sample, threshold, scan (x2 + refine), select
  return n;
}

extern "C" int64_t tt_topk_workspace_bytes(int32_t precision, int64_t nq, int64_t nc, int64_t d, int32_t k) {
  const int kp = pool_k(precision, d, k, nc);
  // [pool scores | pool indices | partial lists of the split scoring stage]
  return round_up(nq * kp * 4, 256) + round_up(nq * kp * 8, 256) + stage1_bytes(precision, nq, nc, d, kp);
}

extern "C" int tt_topk_merge(const float* scores, const int64_t* ids, int32_t num_lists, int64_t nq, int32_t k_in,
                             int32_t k_out, int64_t index_base, const int64_t* identifiers, float* out_scores,
                             int64_t* out_ids, void* stream) {
  TT_REQUIRE(scores && ids && out_scores && out_ids, "tt_topk_merge: null buffer");
  TT_REQUIRE(num_lists >= 1 && num_lists <= 32 * MERGE_LPL, "tt_topk_merge: num_lists must be in [1, %d], got %d", 32 * MERGE_LPL, num_lists);
  TT_REQUIRE(nq >= 0 && k_in >= 1 && k_out >= 1 && k_out <= num_lists * k_in, "tt_topk_merge: bad k (k_in=%d k_out=%d lists=%d)", k_in, k_out, num_lists);
  if (nq == 0) return TT_OK;
  return topk_merge_launch(scores, ids, num_lists, nq, k_in, k_out, index_base, identifiers, out_scores, out_ids, nullptr,
                           (cudaStream_t)stream);
}

static int simt_topk(const void* queries, const void* candidates, int64_t nq, int64_t nc, int64_t d, int k,
                     float* out_scores, int64_t* out_ids, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  TT_REQUIRE(d % 4 == 0 && d <= 256, "tt_topk_bruteforce: fp32 path needs d %% 4 == 0 and d <= 256");
  const size_t smem = simt_topk_smem((int)d, k);
  if (smem > kMaxSmem) return set_error(TT_ERR_UNSUPPORTED, "tt_topk_bruteforce: d=%lld k=%d needs %zu bytes of shared memory (> %zu)", (long long)d, k, smem, kMaxSmem);
  const int splits = simt_num_splits(nq, nc);
  float* ps = out_scores; int64_t* pi = out_ids;
  if (splits > 1) {
    const int64_t need = round_up((int64_t)splits * nq * k * 4, 256) + round_up((int64_t)splits * nq * k * 8, 256);
    if (!workspace || workspace_bytes < need) return set_error(TT_ERR_WORKSPACE, "tt_topk_bruteforce: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need);
    ps = (float*)workspace;
    pi = (int64_t*)((char*)workspace + round_up((int64_t)splits * nq * k * 4, 256));
  }
  const int64_t split_len = round_up(ceil_div(nc, splits), TS);
  dim3 grid((unsigned)ceil_div(nq, TS), (unsigned)splits);
#define TT_TOPK_LAUNCH(KU)                                                                                  \
  {                                                                                                         \
    TT_CUDA_OK(cudaFuncSetAttribute(topk_simt_kernel<KU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    TT_PROF("topk_simt_kernel", st), topk_simt_kernel<KU><<<grid, 256, smem, st>>>((const float*)queries, (const float*)candidates, nq, nc, (int)d, k, \
                                                  0, nullptr, split_len, ps, pi);                           \
  }
  if (k <= 32) TT_TOPK_LAUNCH(1)
  else if (k <= 64) TT_TOPK_LAUNCH(2)
  else if (k <= 128) TT_TOPK_LAUNCH(4)
  else if (k <= 256) TT_TOPK_LAUNCH(8)
  else TT_TOPK_LAUNCH(16)
#undef TT_TOPK_LAUNCH
  TT_LAUNCH_OK("topk_simt_kernel");
  if (splits > 1) return tt_topk_merge(ps, pi, splits, nq, k, k, 0, nullptr, out_scores, out_ids, st);
  return TT_OK;
}

namespace tt {
struct TopkPeerOut {
  unsigned char* const* bases;
  long long off_s, off_i;
  int queries_per_rank, rank;
};
int topk_rerank(int precision, const void* queries, const void* candidates, int64_t nq, int64_t nc, int64_t d,
                const float* pool_s, const int64_t* pool_i, int kp, int k, int64_t base, const int64_t* identifiers,
                float* out_s, int64_t* out_i, int32_t* uncertain, const TopkPeerOut* peer_out, cudaStream_t st);
}

// Two stages: (1) scoring fused with a running top-(k + margin) -- tcgen05 (bf16) or CUDA-core (fp32) -- into a pool
// of raw candidate indices; (2) exact re-rank of the pool (topk_rerank.cu), which fixes the order on the correctly
// rounded fp32 score and applies identifiers / cand_index_base.
static int topk_bruteforce_impl(int32_t precision, const void* queries, const void* candidates, int64_t nq,
                                int64_t nc, int64_t d, int32_t k, int64_t cand_index_base,
                                const int64_t* identifiers, float* out_scores, int64_t* out_ids,
                                int32_t* uncertain_rows, const TopkPeerOut* peer, void* workspace, int64_t workspace_bytes,
                                void* stream) {
  TT_REQUIRE(precision == TT_F32 || precision == TT_BF16, "tt_topk_bruteforce: unknown precision %d", precision);
  TT_REQUIRE(queries && candidates && (peer || (out_scores && out_ids)), "tt_topk_bruteforce: null buffer");
  TT_REQUIRE(nq > 0 && nc > 0 && d > 0 && nc < INT_MAX, "tt_topk_bruteforce: bad sizes");
  TT_REQUIRE(k >= 1 && k <= nc, "tt_topk_bruteforce: k=%d must be in [1, num_candidates=%lld]", k, (long long)nc);
  TT_REQUIRE(k <= 512, "tt_topk_bruteforce: k=%d exceeds the supported maximum 512", k);
  TT_REQUIRE(aligned16(queries) && aligned16(candidates), "tt_topk_bruteforce: inputs must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int kp = pool_k(precision, d, k, nc);
  const int64_t need = tt_topk_workspace_bytes(precision, nq, nc, d, k);
  if (!workspace || workspace_bytes < need)
    return set_error(TT_ERR_WORKSPACE, "tt_topk_bruteforce: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need);
  float* pool_s = (float*)workspace;
  int64_t* pool_i = (int64_t*)((char*)workspace + round_up(nq * kp * 4, 256));
  char* rest = (char*)pool_i + round_up(nq * kp * 8, 256);
  const int64_t rest_bytes = workspace_bytes - (rest - (char*)workspace);
  int rc;
  if (precision == TT_BF16)
    rc = tc_topk(queries, candidates, nq, nc, d, kp, 0, nullptr, pool_s, pool_i, rest, rest_bytes, st);
  else
    rc = simt_topk(queries, candidates, nq, nc, d, kp, pool_s, pool_i, rest, rest_bytes, st);
  if (rc) return rc;
  return topk_rerank(precision, queries, candidates, nq, nc, d, pool_s, pool_i, kp, k, cand_index_base, identifiers,
                     out_scores, out_ids, uncertain_rows, peer, st);
}

// Two stages: (1) scoring fused with a running top-(k + margin) -- tcgen05 (bf16) or CUDA-core (fp32) -- into a pool
// of raw candidate indices; (2) exact re-rank of the pool (topk_rerank.cu), which fixes the order on the correctly
// rounded fp32 score and applies identifiers / cand_index_base.
extern "C" int tt_topk_bruteforce(int32_t precision, const void* queries, const void* candidates, int64_t nq,
                                  int64_t nc, int64_t d, int32_t k, int64_t cand_index_base,
                                  const int64_t* identifiers, float* out_scores, int64_t* out_ids,
                                  int32_t* uncertain_rows, void* workspace, int64_t workspace_bytes, void* stream) {
  return topk_bruteforce_impl(precision, queries, candidates, nq, nc, d, k, cand_index_base, identifiers, out_scores,
                              out_ids, uncertain_rows, nullptr, workspace, workspace_bytes, stream);
}

// Candidate-sharded serving: this rank's shard is scored for ALL nq queries; the exact partial list of query qi is
// written by the re-rank epilogue into list slot [rank] of the receive area of rank qi / queries_per_rank.
extern "C" int tt_topk_bruteforce_peer(int32_t precision, const void* queries, const void* candidates, int64_t nq,
                                       int64_t nc, int64_t d, int32_t k, int64_t cand_index_base,
                                       const int64_t* peer_bases, int32_t world, int32_t rank, int64_t queries_per_rank,
                                       int64_t recv_scores_offset, int64_t recv_ids_offset, int32_t* uncertain_rows,
                                       void* workspace, int64_t workspace_bytes, void* stream) {
  TT_REQUIRE(peer_bases && world >= 1 && world <= 16 && rank >= 0 && rank < world, "tt_topk_bruteforce_peer: bad world / rank");
  TT_REQUIRE(queries_per_rank >= 1 && queries_per_rank * world >= nq, "tt_topk_bruteforce_peer: queries_per_rank * world < nq");
  TT_REQUIRE((recv_scores_offset & 15) == 0 && (recv_ids_offset & 15) == 0, "tt_topk_bruteforce_peer: offsets must be 16-byte aligned");
  TopkPeerOut peer{reinterpret_cast<unsigned char* const*>(peer_bases), recv_scores_offset, recv_ids_offset,
                   (int)queries_per_rank, rank};
  return topk_bruteforce_impl(precision, queries, candidates, nq, nc, d, k, cand_index_base, nullptr, nullptr, nullptr,
                              uncertain_rows, &peer, workspace, workspace_bytes, stream);
}

extern "C" int tt_topk_hits(const float* positive, const float* topk_scores, const int64_t* topk_ids,
                            const int64_t* true_ids, const float* sample_weight, int64_t nq, int32_t k,
                            const int32_t* host_ks, int32_t num_ks, float* hits_out, float* weight_out,
                            void* stream) {
  TT_REQUIRE(host_ks && num_ks >= 1 && num_ks <= 8, "tt_topk_hits: num_ks must be in [1, 8]");
  TT_REQUIRE(hits_out && weight_out, "tt_topk_hits: null output");
  TT_REQUIRE(true_ids ? (topk_ids != nullptr) : (positive && topk_scores), "tt_topk_hits: missing inputs for the chosen mode");
  int ks[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < num_ks; ++i) {
    TT_REQUIRE(host_ks[i] >= 1 && host_ks[i] <= k, "tt_topk_hits: ks[%d]=%d outside [1, k=%d]", i, host_ks[i], k);
    ks[i] = host_ks[i];
  }
  if (nq == 0) return TT_OK;
  TT_PROF("topk_hits_kernel", (cudaStream_t)stream);
  topk_hits_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, (cudaStream_t)stream>>>(
      positive, topk_scores, topk_ids, true_ids, sample_weight, nq, k, nullptr, ks[0], ks[1], ks[2], ks[3], ks[4],
      ks[5], ks[6], ks[7], num_ks, hits_out, weight_out);
  TT_LAUNCH_OK("topk_hits_kernel");
  return TT_OK;
}

extern "C" int tt_rowwise_dot(int32_t precision, const void* q, const void* c, float* out, int64_t n, int64_t d,
                              void* stream) {
  TT_REQUIRE(q && c && out && n >= 0 && d > 0, "tt_rowwise_dot: bad arguments");
  if (n == 0) return TT_OK;
  unsigned blocks = (unsigned)ceil_div(n, 8);
  if (precision == TT_F32) TT_PROF("rowwise_dot_kernel", (cudaStream_t)stream), rowwise_dot_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)q, (const float*)c, out, n, d);
  else if (precision == TT_BF16) TT_PROF("rowwise_dot_kernel", (cudaStream_t)stream), rowwise_dot_kernel<uint16_t><<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)q, (const uint16_t*)c, out, n, d);
  else return set_error(TT_ERR_INVALID_ARG, "tt_rowwise_dot: unknown precision %d", precision);
  TT_LAUNCH_OK("rowwise_dot_kernel");
  return TT_OK;
}
