// K6 (bf16 precision) -- brute-force scoring on tcgen05 fused with a running top-k
// (tfrs BruteForce / faiss.IndexFlatIP replacement).
//
// A CTA keeps 128 queries resident (shared memory, 128B-swizzled K-major) and streams the
// candidate range in 128-row tiles through a TMA ring; S = Q C^T goes to TMEM (two buffers so
// the next tile's MMA overlaps the selection of the current one) and is never written out.
// Selection warpgroup: thread = one query row.  The whole 128-score row of the tile is pulled into registers with
// one tcgen05.ld burst and the TMEM buffer is handed back at once (the next tile's MMA overlaps the selection);
// the thread takes the maximum per 32-score chunk (max3 trees) and compares it with the row's current k-th best
// (tau): almost every chunk is rejected with ~0.4 instructions per score.  A hit is handled warp-cooperatively:
// the chunk is redistributed over the lanes through a 128-byte staging line, survivors are taken one by one.
// The row's k best are kept UNSORTED in shared memory as 64-bit keys (order-preserving score bits << 32 | ~index,
// so one integer compare is the (score desc, index asc) order of tf.math.top_k): an insertion overwrites the
// current minimum and a warp arg-min finds the new one -- no shifting, ~45 instructions.  Rows are sorted once,
// at the end (warp bitonic sort).  Candidate ranges may be split over blockIdx.y for small query batches; partial
// lists are merged by topk_merge_kernel.
#include "tc_common.cuh"
#include "topk_select.cuh"
#include "topk_scan.cuh"
#include <limits.h>

namespace tt {

constexpr int TK_BM = 128;
constexpr int TK_THREADS = 192;   // warp 0 TMA, warp 1 MMA, warps 2-5 selection

struct TopkTcArgs {
  int nq, nc, d, k;
  long long cand_base;
  const long long* identifiers;
  int tiles_per_split;
  int stages;
  int raw_indices;          // partial lists (splits > 1): write raw candidate indices
  const int* run_if;        // optional device flag: the kernel exits at once when it is 0 (fallback of the threshold scan)
  float* out_s;
  long long* out_i;
};

struct TopkTcLayout { int q_bytes, y_bytes, list_bytes, stages, total; };
__host__ __device__ inline TopkTcLayout topk_tc_layout(int d, int k, int TK_BN) {
  TopkTcLayout L;
  L.q_bytes = TK_BM * d * 2;
  L.y_bytes = TK_BN * d * 2;
  L.list_bytes = TK_BM * k * 8;
  const int fixed = L.q_bytes + L.list_bytes + 4096;     // tail: barriers, count/tau, per-warp pending
  L.stages = (227 * 1024 - fixed) / L.y_bytes;
  if (L.stages > 4) L.stages = 4;
  L.total = fixed + L.stages * L.y_bytes;
  return L;
}

template <int KU, int TK_BN>
__global__ void __launch_bounds__(TK_THREADS, 1)
topk_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmC, const TopkTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (a.run_if != nullptr && *a.run_if == 0) return;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int d = a.d, nkb = d / 64, k = a.k;
  const TopkTcLayout L = topk_tc_layout(d, k, TK_BN);
  const int STAGES = a.stages;
  uint8_t* sQ = smem;
  uint8_t* sY = sQ + L.q_bytes;
  uint8_t* tail = sY + STAGES * L.y_bytes;
  uint64_t* q_full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* full = q_full + 1;
  uint64_t* empty = full + 4;
  uint64_t* s_full = empty + 4;
  uint64_t* s_empty = s_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + 2);
  int* row_minpos = reinterpret_cast<int*>(tail + 256);           // [128] slot of the row's current minimum key
  float* stage_x = reinterpret_cast<float*>(tail + 256 + 512);    // [4 warps][32] one chunk of a hit row
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(tail + 4096);   // [128][k], 0 = empty slot

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * TK_BM;
  const int tile_begin = blockIdx.y * a.tiles_per_split;
  const int total_tiles = (a.nc + TK_BN - 1) / TK_BN;
  const int T = max(0, min(a.tiles_per_split, total_tiles - tile_begin));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmC);
    mbar_init(q_full, 1);
    for (int s = 0; s < 4; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&s_full[b], 1); mbar_init(&s_empty[b], 128); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * TK_BN);
  if (threadIdx.x >= 64) row_minpos[threadIdx.x - 64] = 0;
  for (int i = threadIdx.x; i < TK_BM * k; i += TK_THREADS) keys[i] = 0ull;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(q_full, L.q_bytes);
      for (int kb = 0; kb < nkb; ++kb) tma_load_2d(sQ + kb * TK_BM * 128, &tmQ, q_full, kb * 64, q0);
      for (int t = 0; t < T; ++t) {
        const int s = t % STAGES;
        mbar_wait(&empty[s], ((t / STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[s], L.y_bytes);
        for (int kb = 0; kb < nkb; ++kb)
          tma_load_2d(sY + s * L.y_bytes + kb * TK_BN * 128, &tmC, &full[s], kb * 64, (tile_begin + t) * TK_BN);
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc_bf16(TK_BM, TK_BN);
      mbar_wait(q_full, 0);
      for (int t = 0; t < T; ++t) {
        const int s = t % STAGES, b = t & 1;
        mbar_wait(&full[s], (t / STAGES) & 1);
        mbar_wait(&s_empty[b], ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          const uint64_t da = umma_desc_k_sw128(smem_u32(sQ + kb * TK_BM * 128));
          const uint64_t db = umma_desc_k_sw128(smem_u32(sY + s * L.y_bytes + kb * TK_BN * 128));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tmem_base + b * TK_BN, da + 2 * kk, db + 2 * kk, idesc, (kb | kk) != 0);
        }
        umma_commit(&s_full[b]);
        umma_commit(&empty[s]);
      }
    }
  } else {
    const int qd = warp & 3;
    const int wslot = warp - 2;                    // staging line of this warp
    float tau = -INFINITY;                         // k-th best score of this thread's row (-inf until k are held)
    if ((long long)q0 + qd * 32 + lane >= a.nq) tau = INFINITY;          // padding rows never select
    for (int t = 0; t < T; ++t) {
      const int b = t & 1;
      const int c_tile = (tile_begin + t) * TK_BN;
      mbar_wait(&s_full[b], (t >> 1) & 1);
      tc_fence_after();
      uint32_t rr[TK_BN];
#pragma unroll
      for (int c = 0; c < TK_BN / 32; ++c) tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + b * TK_BN + c * 32, rr + c * 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_empty[b]);                    // the MMA of tile t + 2 may overwrite the buffer
      if (c_tile + TK_BN > a.nc) {                 // uniform: ragged last tile
#pragma unroll
        for (int j = 0; j < TK_BN; ++j) if (c_tile + j >= a.nc) rr[j] = 0xff800000u;     // -inf
      }
      float cm[TK_BN / 32];
#pragma unroll
      for (int c = 0; c < TK_BN / 32; ++c) {
        float m0 = fmax3(__uint_as_float(rr[c * 32]), __uint_as_float(rr[c * 32 + 1]), __uint_as_float(rr[c * 32 + 2]));
        float m1 = fmax3(__uint_as_float(rr[c * 32 + 3]), __uint_as_float(rr[c * 32 + 4]), __uint_as_float(rr[c * 32 + 5]));
        float m2 = fmax3(__uint_as_float(rr[c * 32 + 6]), __uint_as_float(rr[c * 32 + 7]), __uint_as_float(rr[c * 32 + 8]));
        float m3 = fmax3(__uint_as_float(rr[c * 32 + 9]), __uint_as_float(rr[c * 32 + 10]), __uint_as_float(rr[c * 32 + 11]));
#pragma unroll
        for (int j = 12; j < 32; j += 10) {        // j = 12, 22: four chains of max3, two scores each
          m0 = fmax3(m0, __uint_as_float(rr[c * 32 + j]), __uint_as_float(rr[c * 32 + j + 1]));
          m1 = fmax3(m1, __uint_as_float(rr[c * 32 + j + 2]), __uint_as_float(rr[c * 32 + j + 3]));
          m2 = fmax3(m2, __uint_as_float(rr[c * 32 + j + 4]), __uint_as_float(rr[c * 32 + j + 5]));
          m3 = fmax3(m3, __uint_as_float(rr[c * 32 + j + 6]), __uint_as_float(rr[c * 32 + j + 7]));
          m0 = fmaxf(m0, __uint_as_float(rr[c * 32 + j + 8]));
          m1 = fmaxf(m1, __uint_as_float(rr[c * 32 + j + 9]));
        }
        cm[c] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      }
      float mx = cm[0];
#pragma unroll
      for (int c = 1; c < TK_BN / 32; ++c) mx = fmaxf(mx, cm[c]);
      unsigned hits = __ballot_sync(0xffffffffu, mx > tau);
      while (hits) {
        const int src = __ffs(hits) - 1;
        hits &= hits - 1;
        const int row = qd * 32 + src;
        unsigned long long* rk = keys + (size_t)row * k;
        float cur_tau = __shfl_sync(0xffffffffu, tau, src);          // uniform; rises as survivors are inserted
#pragma unroll
        for (int c = 0; c < TK_BN / 32; ++c) {
          if (!__shfl_sync(0xffffffffu, (int)(cm[c] > tau), src)) continue;          // uniform
          if (lane == src) {
#pragma unroll
            for (int g = 0; g < 8; ++g)
              *reinterpret_cast<uint4*>(stage_x + wslot * 32 + 4 * g) =
                  make_uint4(rr[c * 32 + 4 * g], rr[c * 32 + 4 * g + 1], rr[c * 32 + 4 * g + 2], rr[c * 32 + 4 * g + 3]);
          }
          __syncwarp();
          const float mine = stage_x[wslot * 32 + lane];              // score (c * 32 + lane) of row `row`
          __syncwarp();
          unsigned pm = __ballot_sync(0xffffffffu, mine > cur_tau);
          while (pm) {
            const int j = __ffs(pm) - 1;
            pm &= pm - 1;
            const float v = __shfl_sync(0xffffffffu, mine, j);
            if (!(v > cur_tau)) continue;                              // tau moved up since the ballot
            cur_tau = topk_replace_min<KU>(rk, row_minpos + row, k, topk_key(v, c_tile + c * 32 + j), lane);
          }
        }
        if (lane == src) tau = cur_tau;
      }
    }
    // sort and write this warp's 32 rows
    __syncwarp();
    for (int rr2 = 0; rr2 < 32; ++rr2) {
      const int row = qd * 32 + rr2;
      const long long qi = (long long)q0 + row;
      if (qi >= a.nq) break;
      unsigned long long v[KU];
#pragma unroll
      for (int u = 0; u < KU; ++u) v[u] = (lane + 32 * u < k) ? keys[(size_t)row * k + lane + 32 * u] : 0ull;
      topk_bitonic_desc<KU>(v, lane);
      const size_t o = ((size_t)blockIdx.y * a.nq + qi) * k;
#pragma unroll
      for (int u = 0; u < KU; ++u) {
        const int t2 = lane + 32 * u;
        if (t2 >= k) continue;
        if (v[u] != 0ull) {
          const int idx = topk_key_index(v[u]);
          a.out_s[o + t2] = topk_key_score(v[u]);
          a.out_i[o + t2] = a.raw_indices ? (long long)idx : (a.identifiers ? a.identifiers[idx] : a.cand_base + idx);
        } else {
          a.out_s[o + t2] = -INFINITY;
          a.out_i[o + t2] = LLONG_MAX;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * TK_BN);
}

static int tk_bn_for(int64_t d) { return d <= 128 ? 128 : 64; }

// splits are planned in 128-candidate granules (independent of the tile width)
static int tc_topk_splits(int64_t nq, int64_t nc, int bn, int* tiles_per_split) {
  const int64_t q_tiles = ceil_div(nq, TK_BM), y128 = ceil_div(nc, 128);
  // Few query tiles (online / small-batch serving): the scan is HBM-bound, so EVERY SM must stream a slice of the
  // candidate matrix -- up to one split per SM (the merge kernel walks up to 256 lists per query).
  int64_t s = std::max<int64_t>(1, num_sms() / q_tiles);
  const int64_t max_by_len = std::max<int64_t>(1, y128 / 32);     // >= 4096 candidates per split
  if (s > max_by_len) s = max_by_len;
  if (s > 256) s = 256;
  const int64_t per128 = ceil_div(y128, s);
  *tiles_per_split = (int)(per128 * (128 / bn));
  return (int)ceil_div(y128, per128);
}

int tc_topk_scan_launches(int64_t nq, int64_t nc, int64_t d, int kp) {
  const ScanPlan p = tc_topk_scan_plan(nq, nc, d, kp);
  return !p.use ? 0 : (p.two_phase ? 6 : 4);
}

// largest k whose key lists leave room for a 2-stage candidate ring
int tc_topk_max_k(int64_t d) {
  if (d % 64 != 0 || d < 64 || d > 256) return 0;
  const int bn = tk_bn_for(d);
  int k = 512;
  while (k > 0 && topk_tc_layout((int)d, k, bn).stages < 2) --k;
  return k;
}

int tc_topk_num_splits(int64_t nq, int64_t nc, int64_t d, int k) {
  int tps;
  return tc_topk_splits(nq, nc, tk_bn_for(d), &tps);
}

int topk_merge_launch(const float* scores, const int64_t* ids, int num_lists, int64_t nq, int k_in, int k_out,
                      int64_t index_base, const int64_t* identifiers, float* out_scores, int64_t* out_ids,
                      const int* run_if, cudaStream_t st);

// workspace of the bf16 scoring stage behind the pool: [partial lists of the list-keeping kernel | threshold scan]
int64_t tc_topk_stage_bytes(int64_t nq, int64_t nc, int64_t d, int kp) {
  int tps;
  const int splits = tc_topk_splits(nq, nc, tk_bn_for(d), &tps);
  int64_t lists = 256;
  if (splits > 1) lists = round_up((int64_t)splits * nq * kp * 4, 256) + round_up((int64_t)splits * nq * kp * 8, 256);
  return lists + tc_topk_scan_plan(nq, nc, d, kp).bytes;
}

int tc_topk(const void* queries, const void* candidates, int64_t nq, int64_t nc, int64_t d, int k,
            int64_t cand_index_base, const int64_t* identifiers, float* out_scores, int64_t* out_ids, void* ws,
            int64_t ws_bytes, cudaStream_t st) {
  TT_REQUIRE(d % 64 == 0 && d >= 64 && d <= 256, "tt_topk_bruteforce(bf16): d must be 64, 128, 192 or 256 (got %lld)", (long long)d);
  const int bn = tk_bn_for(d);
  const TopkTcLayout L = topk_tc_layout((int)d, k, bn);
  if (L.stages < 2)
    return set_error(TT_ERR_UNSUPPORTED, "tt_topk_bruteforce(bf16): d=%lld k=%d does not fit the shared-memory pipeline", (long long)d, k);
  int tps;
  const int splits = tc_topk_splits(nq, nc, bn, &tps);
  float* ps = out_scores; int64_t* pi = out_ids;
  const int64_t lists = splits > 1 ? round_up((int64_t)splits * nq * k * 4, 256) + round_up((int64_t)splits * nq * k * 8, 256) : 256;
  const ScanPlan sp = tc_topk_scan_plan(nq, nc, d, k);
  if (!ws || ws_bytes < lists + sp.bytes)
    return set_error(TT_ERR_WORKSPACE, "tt_topk_bruteforce(bf16): workspace too small (%lld < %lld)", (long long)ws_bytes, (long long)(lists + sp.bytes));
  if (splits > 1) {
    ps = (float*)ws;
    pi = (int64_t*)((char*)ws + round_up((int64_t)splits * nq * k * 4, 256));
  }
  // threshold scan first (topk_scan.cu); the list-keeping kernel below then runs only if a row overflowed
  int* run_if = nullptr;
  if (sp.use) {
    const int rc0 = tc_topk_scan(sp, queries, candidates, nq, nc, d, out_scores, out_ids, (char*)ws + lists, &run_if, st);
    if (rc0) return rc0;
  }
  CUtensorMap tmQ, tmC;
  int rc = make_tmap_bf16_2d(&tmQ, queries, (uint64_t)d, (uint64_t)nq, (uint64_t)d * 2, 64, TK_BM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmC, candidates, (uint64_t)d, (uint64_t)nc, (uint64_t)d * 2, 64, (uint32_t)bn);
  if (rc) return rc;
  TopkTcArgs a{};
  a.nq = (int)nq; a.nc = (int)nc; a.d = (int)d; a.k = k;
  a.cand_base = cand_index_base; a.identifiers = (const long long*)identifiers;
  a.tiles_per_split = tps; a.stages = L.stages; a.raw_indices = splits > 1;
  a.run_if = run_if;
  a.out_s = ps; a.out_i = (long long*)pi;
  dim3 grid((unsigned)ceil_div(nq, TK_BM), (unsigned)splits);
#define TT_TK2(KU, BNV)                                                                                            \
  {                                                                                                                \
    TT_CUDA_OK(cudaFuncSetAttribute(topk_tc_kernel<KU, BNV>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total)); \
    TT_PROF("topk_tc_kernel", st), topk_tc_kernel<KU, BNV><<<grid, TK_THREADS, L.total, st>>>(tmQ, tmC, a);                                       \
  }
#define TT_TK(KU) { if (bn == 128) TT_TK2(KU, 128) else TT_TK2(KU, 64) }
  if (k <= 32) TT_TK(1)
  else if (k <= 64) TT_TK(2)
  else if (k <= 128) TT_TK(4)
  else if (k <= 256) TT_TK(8)
  else TT_TK(16)
#undef TT_TK
#undef TT_TK2
  TT_LAUNCH_OK("topk_tc_kernel");
  if (splits > 1) return topk_merge_launch(ps, pi, splits, nq, k, k, cand_index_base, identifiers, out_scores, out_ids, run_if, st);
  return TT_OK;
}

}  // namespace tt
