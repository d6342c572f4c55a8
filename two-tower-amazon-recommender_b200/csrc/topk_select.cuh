// Per-row running top-k kept in shared memory as a list sorted by (score desc, index asc) --
// the tf.math.top_k order.  Used by the fp32 CUDA-core and the bf16 tcgen05
// scoring kernels: scoring threads filter against the row's current k-th score and append
// survivors to a small pending buffer; one warp then folds the pending entries into the list.
#pragma once
#include "common.cuh"

namespace tt {

struct TopkRowState {
  float* list_s;   // [rows][k]
  int* list_i;     // [rows][k]
  int* count;      // [rows]
  float* tau;      // [rows]   k-th best score once the list is full, else -inf
  float* pend_s;   // [rows][pend_cap]
  int* pend_i;     // [rows][pend_cap]
  int* pend_n;     // [rows]
  int k, pend_cap;
};

__host__ __device__ inline size_t topk_state_bytes(int rows, int k, int pend_cap) {
  return (size_t)rows * ((size_t)k * 8 + (size_t)pend_cap * 8 + 12);
}

__device__ __forceinline__ TopkRowState topk_state_carve(void* base, int rows, int k, int pend_cap) {
  TopkRowState s;
  char* p = (char*)base;
  s.list_s = (float*)p; p += (size_t)rows * k * 4;
  s.list_i = (int*)p;   p += (size_t)rows * k * 4;
  s.pend_s = (float*)p; p += (size_t)rows * pend_cap * 4;
  s.pend_i = (int*)p;   p += (size_t)rows * pend_cap * 4;
  s.count = (int*)p;    p += (size_t)rows * 4;
  s.tau = (float*)p;    p += (size_t)rows * 4;
  s.pend_n = (int*)p;
  s.k = k; s.pend_cap = pend_cap;
  return s;
}

// (s1,i1) strictly precedes (s2,i2) in the output order
__device__ __forceinline__ bool topk_precedes(float s1, int i1, float s2, int i2) {
  return s1 > s2 || (s1 == s2 && i1 < i2);
}

// Warp-cooperative: fold `n` pending (score, index) pairs into one row's sorted list
// (ls/li: [k], *count entries valid).  KU*32 >= k.  Updates *count and *tau.
template <int KU>
__device__ __forceinline__ void topk_fold(float* ls, int* li, int* count, float* tau, int k, const float* pend_s,
                                          const int* pend_i, int n, int lane) {
  int cnt = *count;
  for (int e = 0; e < n; ++e) {
    const float s = pend_s[e];
    const int id = pend_i[e];
    int pos = 0;
    for (int t = lane; t < cnt; t += 32) pos += topk_precedes(ls[t], li[t], s, id) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pos += __shfl_xor_sync(0xffffffffu, pos, o);
    if (pos >= k) continue;
    const int newcnt = min(cnt + 1, k);
    float ts[KU]; int ti[KU];
#pragma unroll
    for (int u = 0; u < KU; ++u) {
      const int t = lane + 32 * u;
      if (t >= pos && t < newcnt - 1) { ts[u] = ls[t]; ti[u] = li[t]; }
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < KU; ++u) {
      const int t = lane + 32 * u;
      if (t >= pos && t < newcnt - 1) { ls[t + 1] = ts[u]; li[t + 1] = ti[u]; }
    }
    if (lane == 0) { ls[pos] = s; li[pos] = id; }
    __syncwarp();
    cnt = newcnt;
  }
  if (lane == 0) {
    *count = cnt;
    *tau = (cnt == k) ? ls[k - 1] : -INFINITY;
  }
  __syncwarp();
}

// Fold row r's own pending buffer (fp32 CUDA-core kernel: one pending buffer per row).
template <int KU>
__device__ __forceinline__ void topk_merge_row(const TopkRowState& st, int r, int lane) {
  const int n = st.pend_n[r];
  if (n == 0) return;
  topk_fold<KU>(st.list_s + (size_t)r * st.k, st.list_i + (size_t)r * st.k, st.count + r, st.tau + r, st.k,
                st.pend_s + (size_t)r * st.pend_cap, st.pend_i + (size_t)r * st.pend_cap, n, lane);
  if (lane == 0) st.pend_n[r] = 0;
  __syncwarp();
}

}  // namespace tt
