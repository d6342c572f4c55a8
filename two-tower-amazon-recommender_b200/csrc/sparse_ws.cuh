// The sparse optimizer's workspace (hash table over the ids of one batch + accumulation rows) and the hash insert
// of its "prepare" stage, shared by sparse_opt.cu and by the fused tower forward (tower_tc.cu), whose control warp
// performs the insert for ID-only towers while the worker warps gather the rows.
#pragma once
#include "common.cuh"
#include <limits.h>

namespace tt {

static constexpr unsigned long long kEmpty = 0xFFFFFFFFFFFFFFFFull;

struct SparseWs {
  unsigned long long* keys;  // [cap]
  int* first;                // [cap]
  int* cnt;                  // [cap] occurrences of the key in this batch (fused step)
  int* done;                 // [cap] arrival tickets of the duplicates (fused step)
  int* hpos;                 // [nnz]
  int* bag_of;               // [nnz]
  float* accum;              // [nnz, d]
  int64_t cap;
};

static inline int64_t hash_capacity(int64_t nnz) {
  int64_t cap = 1024;
  while (cap < 2 * nnz) cap <<= 1;
  return cap;
}

static inline int64_t ws_layout(int64_t nnz, int64_t d, void* base, SparseWs* ws) {
  const int64_t cap = hash_capacity(nnz);
  int64_t off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off += round_up(bytes, 256); return o; };
  int64_t o_keys = take(cap * 8), o_first = take(cap * 4), o_cnt = take(cap * 4), o_done = take(cap * 4),
          o_hpos = take(nnz * 4), o_bag = take(nnz * 4), o_acc = take(nnz * d * 4);
  if (ws) {
    char* b = (char*)base;
    ws->keys = (unsigned long long*)(b + o_keys);
    ws->first = (int*)(b + o_first);
    ws->cnt = (int*)(b + o_cnt);
    ws->done = (int*)(b + o_done);
    ws->hpos = (int*)(b + o_hpos);
    ws->bag_of = (int*)(b + o_bag);
    ws->accum = (float*)(b + o_acc);
    ws->cap = cap;
  }
  return off;
}

// The layout of a caller-owned workspace is a function of its SIZE, not of the batch at hand: a workspace made for
// nnz entries is reused for every batch with at most that many (ragged bags change nnz from step to step), and every
// array boundary -- in particular the accumulation rows, which must be zero between launches -- has to stay where
// tt_sparse_workspace_init put it.  capacity = the largest entry count whose layout fits the buffer.
static inline int64_t ws_capacity(int64_t workspace_bytes, int64_t d) {
  if (ws_layout(0, d, nullptr, nullptr) > workspace_bytes) return -1;
  int64_t lo = 0, hi = 1;
  while (hi < ((int64_t)1 << 31) && ws_layout(hi, d, nullptr, nullptr) <= workspace_bytes) { lo = hi; hi <<= 1; }
  while (hi - lo > 1) {
    const int64_t mid = lo + (hi - lo) / 2;
    if (ws_layout(mid, d, nullptr, nullptr) <= workspace_bytes) lo = mid; else hi = mid;
  }
  return lo;
}
// carve `workspace` for a batch of nnz entries; false if it is too small
static inline bool ws_carve(int64_t nnz, int64_t d, void* workspace, int64_t workspace_bytes, SparseWs* ws) {
  const int64_t cap = ws_capacity(workspace_bytes, d);
  if (cap < nnz) return false;
  ws_layout(cap, d, workspace, ws);
  return true;
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}


// Prepare stage for U entries of one thread: entry j[u] holds the (validated, rank-local) row id[u]; pend[u] marks the
// live ones.  Finds / claims the key's slot (linear probing from mix64(id)), records it in hpos[j] and folds j into the
// slot's first occurrence and count.
// A probe is a round trip to L2 (~1 us under load) and the table runs at a load factor of up to 0.5, where the longest
// of 8192 linear-probe chains is 15-20 slots: probing slot by slot with atomicCAS made the whole stage 12-16 us.  Here
// a probe reads the 64-byte line of 8 keys around the position (L1-bypassing loads), picks the first slot at or behind
// it that is empty or already holds the key, and only then issues ONE atomicCAS (none if the key is there): typically
// two round trips per entry, and a thread keeps the round trips of its U entries in flight together.  Slots only ever
// go from empty to occupied during this stage, so "first empty-or-equal slot in probe order, confirmed by the CAS" picks
// the same slot for every inserter of a key, exactly as slot-by-slot probing does.
template <int U>
__device__ __forceinline__ void sparse_prepare_entries(const SparseWs& ws, const int64_t (&j)[U], const int64_t (&id)[U],
                                                       bool (&pend)[U]) {
  const uint64_t mask = (uint64_t)ws.cap - 1;      // cap: a power of two >= 1024, so a line of 8 keys never wraps
  uint64_t pos[U];
  bool any = false;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    pos[u] = mix64((uint64_t)id[u]) & mask;
    any |= pend[u];
  }
  while (any) {
    ulonglong2 w[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (pend[u]) {
        const ulonglong2* line = reinterpret_cast<const ulonglong2*>(ws.keys + (pos[u] & ~7ull));
#pragma unroll
        for (int k = 0; k < 4; ++k) w[u][k] = __ldcg(line + k);
      }
    int cand[U];
    bool hit[U];
    unsigned long long prev[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      cand[u] = -1;
      hit[u] = false;
      prev[u] = 0;
      if (!pend[u]) continue;
      const int s0 = (int)(pos[u] & 7);
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const unsigned long long key = (s & 1) ? w[u][s >> 1].y : w[u][s >> 1].x;
        if (cand[u] < 0 && s >= s0 && (key == kEmpty || key == (unsigned long long)id[u])) {
          cand[u] = s;
          hit[u] = key == (unsigned long long)id[u];
        }
      }
      if (cand[u] >= 0 && !hit[u])
        prev[u] = atomicCAS(&ws.keys[(pos[u] & ~7ull) + cand[u]], kEmpty, (unsigned long long)id[u]);
    }
    any = false;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!pend[u]) continue;
      if (cand[u] < 0) {                            // the rest of the line belongs to other keys
        pos[u] = ((pos[u] & ~7ull) + 8) & mask;
        any = true;
        continue;
      }
      const uint64_t slot = (pos[u] & ~7ull) + cand[u];
      if (hit[u] || prev[u] == kEmpty || prev[u] == (unsigned long long)id[u]) {
        ws.hpos[j[u]] = (int)slot;
        atomicMin(&ws.first[slot], (int)j[u]);
        atomicAdd(&ws.cnt[slot], 1);
        pend[u] = false;
      } else {                                      // another key won the slot in between: go on behind it
        pos[u] = (slot + 1) & mask;
        any = true;
      }
    }
  }
}

__device__ __forceinline__ void sparse_prepare_entry(const SparseWs& ws, int64_t j, int64_t id) {
  const int64_t jj[1] = {j}, ii[1] = {id};
  bool pend[1] = {true};
  sparse_prepare_entries<1>(ws, jj, ii, pend);
}

}  // namespace tt
