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

// ---- unsorted k-best as 64-bit keys (tcgen05 scoring kernel) ---------------------------------------------
// key = order-preserving bits of the score << 32 | ~index: a larger key precedes in the tf.math.top_k order
// (score desc, index asc); 0 marks an empty slot (below every real key).
__device__ __forceinline__ unsigned long long topk_key(float s, int idx) {
  const unsigned u = __float_as_uint(s);
  const unsigned f = u ^ ((unsigned)((int)u >> 31) | 0x80000000u);
  return ((unsigned long long)f << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)idx);
}
__device__ __forceinline__ float topk_key_score(unsigned long long key) {
  const unsigned f = (unsigned)(key >> 32);
  return __uint_as_float((f & 0x80000000u) ? (f ^ 0x80000000u) : ~f);
}
__device__ __forceinline__ int topk_key_index(unsigned long long key) { return (int)(0xFFFFFFFFu - (unsigned)key); }

// Warp-cooperative: overwrite the row's minimum key (slot *minpos) with `key` (the caller checked key > minimum),
// find the new minimum.  Returns the new k-th best score (-inf while empty slots remain).  KU * 32 >= k.
template <int KU>
__device__ __forceinline__ float topk_replace_min(unsigned long long* rk, int* minpos, int k, unsigned long long key, int lane) {
  if (lane == 0) rk[*minpos] = key;
  __syncwarp();
  unsigned long long lmin = ~0ull;
  int lpos = 0;
#pragma unroll
  for (int u = 0; u < KU; ++u) {
    const int t = lane + 32 * u;
    if (t < k) {
      const unsigned long long kk = rk[t];
      if (kk < lmin) { lmin = kk; lpos = t; }        // ascending t: the lowest slot wins among equal (empty) keys
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long om = __shfl_xor_sync(0xffffffffu, lmin, o);
    const int op = __shfl_xor_sync(0xffffffffu, lpos, o);
    if (om < lmin || (om == lmin && op < lpos)) { lmin = om; lpos = op; }
  }
  if (lane == 0) *minpos = lpos;
  __syncwarp();
  return lmin == 0ull ? -INFINITY : topk_key_score(lmin);
}

// Warp bitonic sort, descending, of KU * 32 keys: element e = u * 32 + lane lives in v[u] of lane `lane`.
template <int KU>
__device__ __forceinline__ void topk_bitonic_desc(unsigned long long (&v)[KU], int lane) {
#pragma unroll
  for (int size = 2; size <= KU * 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= 32) {
        const int du = stride / 32;
#pragma unroll
        for (int u = 0; u < KU; ++u) {
          if (u & du) continue;
          const bool desc = (((u * 32 + lane) & size) == 0);
          const unsigned long long lo = v[u] < v[u + du] ? v[u] : v[u + du], hi = v[u] < v[u + du] ? v[u + du] : v[u];
          v[u] = desc ? hi : lo;
          v[u + du] = desc ? lo : hi;
        }
      } else {
#pragma unroll
        for (int u = 0; u < KU; ++u) {
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, v[u], stride);
          const bool desc = (((u * 32 + lane) & size) == 0);
          const bool first = (lane & stride) == 0;               // the lower element of the pair
          const bool want_max = first == desc;
          v[u] = want_max ? (v[u] > other ? v[u] : other) : (v[u] < other ? v[u] : other);
        }
      }
    }
  }
}

}  // namespace tt
