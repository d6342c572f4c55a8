// K6 (bf16 precision) -- threshold scan: brute-force scoring on tcgen05 where the running top-k of topk_tc.cu is
// replaced by a per-row score THRESHOLD known before the scan starts (tfrs BruteForce / faiss.IndexFlatIP
// replacement; BASELINE configs[4]).
//
// A running top-k pays k (1 + ln(n / k)) list insertions per query row and candidate split -- with one split per SM
// (small query batches, HBM-bound) that is ~1 % of ALL scores, each a warp-cooperative update: 12 ms for a scan that
// moves 2.56 GB.  Here no list is kept while scanning:
//   1. sample    (topk_scan_kernel<.., true>): score an evenly strided ~1/32 of the candidate tiles and write the row
//                maximum of every 32-candidate group;
//   2. threshold (topk_tau_kernel): tau[row] = the K-th largest of the row's G group maxima.  The K maxima above it
//                are K distinct candidates, so AT LEAST K candidates score >= tau -- a guarantee, not an estimate --
//                and about K * (n / 32) / G of all n do;
//   3. scan      (topk_scan_kernel<.., false>): every SM streams candidate tiles (TMA ring -> tcgen05 -> TMEM); the
//                selection threads (thread = query row, the 128 scores of a tile in registers) compare 16-score maxima
//                with tau -- 0.5 instructions per score, no shared state -- and PARK flagged 16-score groups in a
//                thread-private shared-memory FIFO (predicated stores, all lanes at once); a rolled, branch-free drain
//                moves the survivors as 64-bit keys into the segment of the row's survivor buffer that belongs to this
//                CTA alone (one segment per row and candidate split: no atomics, the write position is a register).
//                Tiles are dealt to the splits round-robin, so that candidates stored in an order that correlates with
//                the scores still spread evenly over the segments;
//   4. select    (topk_pool_select_kernel): one CTA per row gathers the row's ~1-4 K survivors from its segments, radix-
//                selects the K-th largest key and sorts the K keys at or above it ((score desc, index asc) order) -- the
//                pool the exact re-rank consumes.
// More than 128 resident queries: the scan is bound by the tensor pipe and the selection warps, not by HBM, so every
// survivor costs issue slots: an eighth of the tiles is scanned first, topk_tau_refine_kernel raises each row's threshold
// to the K-th largest survivor so far (still a guarantee), and the remaining tiles see 3x fewer survivors.
// The scores of steps 1 and 3 are the same MMA chain on the same operands (bit-identical), so step 4 finds at least K
// survivors unless a segment overflows (adversarial data); then a device-side flag makes the list-keeping kernel of
// topk_tc.cu run instead (launched behind the flag, it exits at once otherwise).
//
// CTA shapes.  <= 128 queries: one query tile in TENSOR MEMORY (TS-form MMA: no shared-memory reads for A), three
// rotating accumulator buffers, its shared-memory staging aliased onto the ring's last stage -> a 6-tile ring; 64 queries
// x 10 M candidates stream at 6.4 TB/s, 0.98 of the measured copy rate.  > 128 queries: two query tiles per CTA share
// every candidate tile (half the L2 -> shared-memory traffic per flop), SS-form MMAs, two PRIVATE accumulator buffers
// and ONE MMA-ISSUING THREAD PER QUERY TILE: a single issuer spent ~250 cycles in each of three mbarrier waits and ~600
// issuing 16 MMAs per tile -- 1740 cycles against 1024 of tensor-pipe work (tools/trace_topk_scan.py).
#include "tc_common.cuh"
#include "topk_select.cuh"
#include "topk_scan.cuh"
#include <limits.h>
#include <algorithm>
#include <stdlib.h>

namespace tt {

constexpr int SC_BM = 128, SC_BN = 128;
// Thread-private FIFO of flagged 16-score groups in shared memory: entry = 16 scores + the index of the first one.
// Layout [slot][4 x 16 B][thread] (consecutive threads 16 B apart: conflict-free STS.128) + [slot][thread] indices.
__host__ __device__ constexpr int sc_depth(int nqt) { return nqt == 1 ? 4 : 3; }
constexpr int SC_MAX_STAGES = 6;
constexpr int SC_SBUF = 3;             // accumulator buffers in tensor memory: columns [0, 384)
constexpr uint32_t SC_Q_COL = 384;     // query tiles as packed bf16x2: 64 columns each (d <= 128)

struct ScanArgs {
  int nq, nc, d;
  int total_tiles;          // ceil(nc / 128)
  int tiles_per_split;      // scan: contiguous tiles per blockIdx.y; sample: sample tiles per blockIdx.y
  int n_samp;               // sample: number of sampled tiles (tile of sample i = i * total_tiles / n_samp)
  int phase;                // scan: 0 = every tile; 1 = the tiles = 0 mod 8; 2 = the other tiles (after the threshold was refined)
  int nq_pad, stages;
  int n_slots, slot0, seg_cap;   // scan: survivor segments per row, first segment of this launch, entries per segment
  const float* tau;         // scan: [nq_pad]
  int* cnt;                 // scan: [nq][n_slots] survivors per (row, segment) (may exceed seg_cap: overflow)
  unsigned long long* buf;  // scan: [nq][n_slots][seg_cap]
  float* samp;              // sample: [nq][4 * n_samp]
  long long* trace;         // debug (tt_debug_topk_scan_trace): clock64() stamps of CTA (0,0), [role][tile][4], else null
};

constexpr int SC_TRACE_TILES = 48;
static long long* g_scan_trace = nullptr;
#define SC_TRACE(role, tile, ev)                                                              \
  do {                                                                                        \
    if (a.trace && blockIdx.x == 0 && blockIdx.y == 0 && (tile) < SC_TRACE_TILES)             \
      a.trace[((role) * SC_TRACE_TILES + (tile)) * 4 + (ev)] = clock64();                     \
  } while (0)

struct ScanLayout { int q_bytes, y_bytes, pq_bytes, stages, total; };
__host__ __device__ inline ScanLayout scan_layout(int d, int nqt, bool sample) {
  ScanLayout L;
  L.q_bytes = nqt * SC_BM * d * 2;
  L.y_bytes = SC_BN * d * 2;
  L.pq_bytes = sample ? 0 : nqt * SC_BM * sc_depth(nqt) * (64 + 4);
  // One resident query tile (TS form): the tile only passes through shared memory on its way to tensor memory, so
  // it ALIASES the last stage of the candidate ring (the producer touches that stage only after the copy).
  // Two resident query tiles (SS form): they stay in shared memory next to the ring.
  const int fixed = L.pq_bytes + 1024 + (nqt == 1 ? 0 : L.q_bytes);
  L.stages = (227 * 1024 - fixed) / L.y_bytes;
  if (L.stages > SC_MAX_STAGES) L.stages = SC_MAX_STAGES;
  if (nqt == 1 && L.stages * L.y_bytes < L.q_bytes + 2 * L.y_bytes) L.stages = 0;      // needs two stages of its own to start
  L.total = fixed + L.stages * L.y_bytes;
  return L;
}

// Drain one thread's FIFO into the segment of the row's survivor buffer that belongs to THIS CTA alone (one segment
// per (row, candidate split): no atomics, no counting pass, the write position lives in a register of the row's thread).
// Rolled loop: this runs once per ~10 tiles, lanes in parallel.
__device__ __noinline__ int scan_drain(const uint4* fifo, const int* loc, int n, float tau, unsigned long long* seg,
                                       int seg_cap, int pos) {
#pragma unroll 1
  for (int e = 0; e < 4 * n; ++e) {
    const uint4 v = fifo[e * SC_BM];
    const int i0 = loc[(e >> 2) * SC_BM] + 4 * (e & 3);
    const uint32_t x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {                  // branch-free: the key is always formed, the store is predicated
      const bool in = __uint_as_float(x[u]) > tau;
      const unsigned long long key = topk_key(__uint_as_float(x[u]), i0 + u);
      if (in && pos < seg_cap) seg[pos] = key;
      pos += in ? 1 : 0;
    }
  }
  return pos;
}

template <int NQT, bool SAMPLE>
__global__ void __launch_bounds__(32 * (1 + NQT) + NQT * 128, 1)
topk_scan_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmC, const ScanArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int d = a.d, nkb = d / 64;
  const ScanLayout L = scan_layout(d, NQT, SAMPLE);
  const int STAGES = a.stages;
  // NQT == 1: TS form -- the query tile lives in tensor memory, three accumulator buffers rotate, one MMA issuer.
  // NQT == 2: SS form -- query tiles in shared memory, two PRIVATE accumulator buffers and one MMA issuer per query tile
  // (shared rotating buffers would let one issuer run two barrier phases ahead of the other tile's selection warps,
  // which a parity wait cannot tell from zero phases ahead).
  constexpr bool TS = NQT == 1;
  uint8_t* sY = smem;
  uint8_t* sQ = TS ? sY + STAGES * L.y_bytes - L.q_bytes : sY + STAGES * L.y_bytes;   // TS: aliases the ring's tail until q_ready
  unsigned long long* pq_all = reinterpret_cast<unsigned long long*>(sY + STAGES * L.y_bytes + (TS ? 0 : L.q_bytes));
  uint8_t* tail = reinterpret_cast<uint8_t*>(pq_all) + L.pq_bytes;
  uint64_t* q_full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* full = q_full + 1;                    // [SC_MAX_STAGES]
  uint64_t* empty = full + SC_MAX_STAGES;         // [SC_MAX_STAGES]
  uint64_t* s_full = empty + SC_MAX_STAGES;       // [4] accumulator buffers
  uint64_t* s_empty = s_full + 4;                 // [4]
  uint64_t* q_ready = s_empty + 4;                // TS: the query tile sits in tensor memory
  // accumulator buffer and use count of (tile t, query tile j)
  auto buf_of = [&](int t, int j) -> int { return TS ? t % SC_SBUF : 2 * j + (t & 1); };
  auto use_of = [&](int t) -> int { return TS ? t / SC_SBUF : t >> 1; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_ready + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (NQT * SC_BM);
  // sample: a contiguous run of sampled tiles per split.  scan: tile slots blockIdx.y, + gridDim.y, ... -- strided, so
  // that candidates ordered by anything that correlates with the scores still spread evenly over the splits' segments
  const int begin = SAMPLE ? blockIdx.y * a.tiles_per_split : blockIdx.y;
  const int step = SAMPLE ? 1 : gridDim.y;
  const int n_first = (a.total_tiles + 7) >> 3;                     // tiles = 0 mod 8
  const int limit = SAMPLE ? a.n_samp : (a.phase == 0 ? a.total_tiles : (a.phase == 1 ? n_first : a.total_tiles - n_first));
  const int T = SAMPLE ? max(0, min(a.tiles_per_split, limit - begin)) : (limit > begin ? (limit - begin + step - 1) / step : 0);
  auto tile_of = [&](int t) -> int {
    const int i = begin + t * step;
    if (SAMPLE) return (int)(((long long)i * a.total_tiles) / a.n_samp);
    if (a.phase == 0) return i;
    if (a.phase == 1) return i << 3;
    const int q7 = i / 7;
    return (q7 << 3) + (i - 7 * q7) + 1;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmC);
    mbar_init(q_full, 1);
    for (int s = 0; s < SC_MAX_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NQT); }
    for (int b = 0; b < 4; ++b) { mbar_init(&s_full[b], 1); mbar_init(&s_empty[b], 128); }
    mbar_init(q_ready, NQT * 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(q_full, L.q_bytes);
      for (int j = 0; j < NQT; ++j)
        for (int kb = 0; kb < nkb; ++kb)
          tma_load_2d(sQ + (j * nkb + kb) * SC_BM * 128, &tmQ, q_full, kb * 64, q0 + j * SC_BM);
      for (int t = 0; t < T; ++t) {
        const int s = t % STAGES;
        if (TS && t < STAGES && (s + 1) * L.y_bytes > STAGES * L.y_bytes - L.q_bytes) mbar_wait(q_ready, 0);   // stage under the query tile
        mbar_wait(&empty[s], ((t / STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[s], L.y_bytes);
        const int y0 = tile_of(t) * SC_BN;
        for (int kb = 0; kb < nkb; ++kb) tma_load_2d(sY + s * L.y_bytes + kb * SC_BN * 128, &tmC, &full[s], kb * 64, y0);
      }
    }
  } else if (warp <= NQT) {
    // One MMA-issuing thread PER QUERY TILE (warps 1 .. NQT).  A single issuer spends ~250 cycles in each of its
    // three mbarrier waits per candidate tile and ~300 issuing eight MMAs with their descriptor arithmetic -- 1740
    // cycles per tile against 1024 of tensor-pipe work (tools/trace_topk_scan.py); two issuers overlap each other's waits.
    if (elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc_bf16(SC_BM, SC_BN);
      const int j = warp - 1;
      if (TS) {
        mbar_wait(q_ready, 0);
        tc_fence_after();
      } else {
        mbar_wait(q_full, 0);
      }
      for (int t = 0; t < T; ++t) {
        const int s = t % STAGES;
        const int b = buf_of(t, j), k = use_of(t);
        if (j == 0) SC_TRACE(2, t, 0);
        mbar_wait(&full[s], (t / STAGES) & 1);
        if (j == 0) SC_TRACE(2, t, 1);
        mbar_wait(&s_empty[b], (k & 1) ^ 1);
        if (j == 0) SC_TRACE(2, t, 2);
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          const uint64_t db = umma_desc_k_sw128(smem_u32(sY + s * L.y_bytes + kb * SC_BN * 128));
          if (TS) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16_ts(tmem_base + b * SC_BN, tmem_base + SC_Q_COL + 8 * (4 * kb + kk), db + 2 * kk, idesc, (kb | kk) != 0);
          } else {
            const uint64_t da = umma_desc_k_sw128(smem_u32(sQ + (j * nkb + kb) * SC_BM * 128));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16_ss(tmem_base + b * SC_BN, da + 2 * kk, db + 2 * kk, idesc, (kb | kk) != 0);
          }
        }
        umma_commit(&s_full[b]);
        umma_commit(&empty[s]);                    // NQT arrivals release the stage
        if (j == 0) SC_TRACE(2, t, 3);
      }
    }
  } else {
    const int j = (warp - 1 - NQT) >> 2;           // query tile of this warpgroup
    const int qd = warp & 3;                       // TMEM lane quadrant this warp may read
    const int row = qd * 32 + lane;
    const long long qi = (long long)q0 + j * SC_BM + row;
    float tau = INFINITY;                          // padding rows never select
    constexpr int DEPTH = sc_depth(NQT);
    uint4* fifo = nullptr;                         // this thread's FIFO of flagged 16-score groups
    int* loc = nullptr;
    int nfifo = 0;
    unsigned long long* seg = nullptr;             // this thread's (row, split) segment of the survivor buffer
    int pos = 0;
    if constexpr (!SAMPLE) {
      if (qi < a.nq) {
        tau = a.tau[qi];
        seg = a.buf + ((size_t)qi * a.n_slots + a.slot0 + blockIdx.y) * a.seg_cap;
      }
      fifo = reinterpret_cast<uint4*>(pq_all) + (size_t)j * DEPTH * 4 * SC_BM + row;
      loc = reinterpret_cast<int*>(reinterpret_cast<uint4*>(pq_all) + (size_t)NQT * DEPTH * 4 * SC_BM) + (size_t)j * DEPTH * SC_BM + row;
    }
    if constexpr (TS) {
      // the stationary query tile: shared memory (TMA, zero-filled past nq) -> registers -> tensor memory, row r by
      // thread r as packed bf16x2 (one 32-bit column = two consecutive K elements, what the TS-form MMA reads)
      mbar_wait(q_full, 0);
      for (int kb = 0; kb < nkb; ++kb) {
        uint32_t qq[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 v = *reinterpret_cast<const uint4*>(sQ + (j * nkb + kb) * SC_BM * 128 + sw128_offset(row, c));
          qq[4 * c] = v.x; qq[4 * c + 1] = v.y; qq[4 * c + 2] = v.z; qq[4 * c + 3] = v.w;
        }
        tmem_st32(tmem_base + ((uint32_t)(qd * 32) << 16) + SC_Q_COL + j * 64 + kb * 32, qq);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(q_ready);
    }
    for (int t = 0; t < T; ++t) {
      const int b = buf_of(t, j);
      const int c_tile = tile_of(t) * SC_BN;
      if (row == 0) SC_TRACE(j, t, 0);
      mbar_wait(&s_full[b], use_of(t) & 1);
      if (row == 0) SC_TRACE(j, t, 1);
      tc_fence_after();
      uint32_t rr[SC_BN];
#pragma unroll
      for (int c = 0; c < SC_BN / 32; ++c)
        tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + b * SC_BN + c * 32, rr + c * 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_empty[b]);                    // the buffer's next unit may overwrite it
      if (row == 0) SC_TRACE(j, t, 2);
      if (c_tile + SC_BN > a.nc) {                 // uniform: ragged last tile
#pragma unroll
        for (int i = 0; i < SC_BN; ++i) if (c_tile + i >= a.nc) rr[i] = 0xff800000u;     // -inf
      }
      float sm[SC_BN / 16];                        // maxima of 16 scores
#pragma unroll
      for (int c = 0; c < SC_BN / 16; ++c) {
        const uint32_t* x = rr + c * 16;
        float m0 = fmax3(__uint_as_float(x[0]), __uint_as_float(x[1]), __uint_as_float(x[2]));
        float m1 = fmax3(__uint_as_float(x[3]), __uint_as_float(x[4]), __uint_as_float(x[5]));
        m0 = fmax3(m0, __uint_as_float(x[6]), __uint_as_float(x[7]));
        m1 = fmax3(m1, __uint_as_float(x[8]), __uint_as_float(x[9]));
        m0 = fmax3(m0, __uint_as_float(x[10]), __uint_as_float(x[11]));
        m1 = fmax3(m1, __uint_as_float(x[12]), __uint_as_float(x[13]));
        m0 = fmax3(m0, __uint_as_float(x[14]), __uint_as_float(x[15]));
        sm[c] = fmaxf(m0, m1);
      }
      if constexpr (SAMPLE) {
        if (qi < a.nq)                               // 16 bytes per row and sampled tile: the threshold kernel reads rows
          *reinterpret_cast<float4*>(a.samp + (size_t)qi * (4 * a.n_samp) + (size_t)(begin + t) * 4) =
              make_float4(fmaxf(sm[0], sm[1]), fmaxf(sm[2], sm[3]), fmaxf(sm[4], sm[5]), fmaxf(sm[6], sm[7]));
      } else {
        float mx = sm[0];
#pragma unroll
        for (int c = 1; c < SC_BN / 16; ++c) mx = fmaxf(mx, sm[c]);
        // Flagged groups (a few % of the row-tiles hold one) are only PARKED here -- predicated stores, every lane at
        // once -- and examined later by scan_drain.  Examining them in place costs either 40 KB of unrolled code (the
        // kernel then stalls on instruction fetch) or a serial dependent loop per hit row with one warp per scheduler.
        if (__any_sync(0xffffffffu, mx > tau)) {
#pragma unroll
          for (int c = 0; c < SC_BN / 16; ++c) {
            if (sm[c] > tau) {
              if (nfifo == DEPTH) {               // one row with a burst: drained alone
                pos = scan_drain(fifo, loc, nfifo, tau, seg, a.seg_cap, pos);
                nfifo = 0;
              }
#pragma unroll
              for (int g = 0; g < 4; ++g)
                fifo[(nfifo * 4 + g) * SC_BM] = make_uint4(rr[c * 16 + 4 * g], rr[c * 16 + 4 * g + 1], rr[c * 16 + 4 * g + 2], rr[c * 16 + 4 * g + 3]);
              loc[nfifo * SC_BM] = c_tile + c * 16;
              ++nfifo;
            }
          }
          if (__any_sync(0xffffffffu, nfifo == DEPTH)) {      // the whole warp drains together
            if (nfifo > 0) pos = scan_drain(fifo, loc, nfifo, tau, seg, a.seg_cap, pos);
            nfifo = 0;
          }
        }
      }
      if (row == 0) SC_TRACE(j, t, 3);
    }
    if constexpr (!SAMPLE) {
      if (nfifo > 0) pos = scan_drain(fifo, loc, nfifo, tau, seg, a.seg_cap, pos);
      if (qi < a.nq) a.cnt[(size_t)qi * a.n_slots + a.slot0 + blockIdx.y] = pos;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---- block-wide bitonic sort (descending) of P = 2^m keys in shared memory -----------------------------------
template <typename K>
__device__ __forceinline__ void block_bitonic_desc(K* s, int P) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < (P >> 1); i += blockDim.x) {
        const int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
        const int hi = lo | stride;
        const bool desc = (lo & size) == 0;
        const K x = s[lo], y = s[hi];
        if ((x < y) == desc) { s[lo] = y; s[hi] = x; }
      }
      __syncthreads();
    }
  }
}

// ---- block-wide radix select: the k-th largest (k >= 1) of the keys the block holds in registers ----------------
// Key 0 is padding (below every real key; k never exceeds the number of real keys).  Eight bits per pass from the
// top; equal digits inside a warp are combined before the shared-memory atomic (scores of one row share their high
// bits, so the plain form would serialise 32 ways).  hist: shared unsigned[257], bc: shared int[2].
template <typename K, int NPT>
__device__ __forceinline__ K block_kth_largest(const K (&key)[NPT], int n_valid, int k, unsigned* hist, int* bc) {
  const int lane = threadIdx.x & 31;
  K prefix = 0, mask = 0;
  for (int shift = (int)sizeof(K) * 8 - 8; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 257; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      if (i * (int)blockDim.x >= n_valid) break;           // uniform: the rest is padding
      const unsigned dg = (key[i] & mask) == prefix ? (unsigned)((key[i] >> shift) & 255) : 256u;
      const unsigned peers = __match_any_sync(0xffffffffu, dg);
      if (lane == __ffs(peers) - 1) atomicAdd(&hist[dg], (unsigned)__popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < 32) {                       // lane l owns bins 255 - 8 l ... 248 - 8 l, highest first
      unsigned c[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { c[j] = hist[255 - 8 * lane - j]; sum += c[j]; }
      unsigned incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      unsigned run = incl - sum;
      if (run < (unsigned)k && (unsigned)k <= incl) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (run < (unsigned)k && (unsigned)k <= run + c[j]) { bc[0] = 255 - 8 * lane - j; bc[1] = k - (int)run; }
          run += c[j];
        }
      }
    }
    __syncthreads();
    prefix |= (K)(unsigned)bc[0] << shift;
    mask |= (K)255 << shift;
    k = bc[1];
  }
  return prefix;
}

__device__ __forceinline__ unsigned score_key32(float s) {
  const unsigned u = __float_as_uint(s);
  return u ^ ((unsigned)((int)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ float key32_score(unsigned f) {
  return __uint_as_float((f & 0x80000000u) ? (f ^ 0x80000000u) : ~f);
}

constexpr int SC_SEL_THREADS = 1024;
constexpr int SC_TAU_NPT = 32;         // G <= 32768 group maxima per row
constexpr int SC_SEL_NPT = 16;         // cap <= 16384 survivors per row

// tau[row] = the largest float strictly below the K-th largest of the row's G group maxima (so `score > tau` means
// `score >= that maximum`); -inf when fewer than K groups hold a finite score.  Also clears the row's survivor count.
__device__ __forceinline__ float just_below(float kth) {
  if (!(kth > -INFINITY)) return -INFINITY;                            // also NaN
  if (kth == 0.f) return __uint_as_float(0x80000001u);                 // below both zeros
  return key32_score(score_key32(kth) - 1u);
}

__global__ void __launch_bounds__(SC_SEL_THREADS)
topk_tau_kernel(const float* __restrict__ samp, int G, int K, float* __restrict__ tau, int* __restrict__ cnt,
                int n_slots, int* __restrict__ flag) {
  __shared__ unsigned hist[257];
  __shared__ int bc[2];
  const int row = blockIdx.x;
  unsigned key[SC_TAU_NPT];
#pragma unroll
  for (int i = 0; i < SC_TAU_NPT; ++i) {
    const int g = threadIdx.x + i * SC_SEL_THREADS;
    key[i] = g < G ? score_key32(samp[(size_t)row * G + g]) : 0u;
  }
  for (int i = threadIdx.x; i < n_slots; i += SC_SEL_THREADS) cnt[(size_t)row * n_slots + i] = 0;
  const unsigned kk = block_kth_largest<unsigned, SC_TAU_NPT>(key, G, K, hist, bc);
  if (threadIdx.x == 0) {
    tau[row] = just_below(key32_score(kk));
    if (row == 0) *flag = 0;
  }
}

// The survivors of one row sit in n_slots segments of seg_cap entries.  Prefix sums of the segment counts in shared
// memory (pre[0 .. n_slots]); returns the total, or -1 if a segment overflowed.  n_slots <= SC_MAX_SLOTS.
constexpr int SC_MAX_SLOTS = 1024;
__device__ __forceinline__ int segment_prefix(const int* __restrict__ cnt_row, int n_slots, int seg_cap, int* pre, int* flag_s) {
  if (threadIdx.x == 0) *flag_s = 0;
  __syncthreads();
  int c = 0;
  if ((int)threadIdx.x < n_slots) {
    c = cnt_row[threadIdx.x];
    if (c > seg_cap) { *flag_s = 1; c = seg_cap; }
  }
  // inclusive scan over the first n_slots threads (n_slots <= blockDim.x = 1024): warp scan + warp totals
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int v = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += u;
  }
  __shared__ int wsum[32];
  if (lane == 31) wsum[w] = v;
  __syncthreads();
  if (w == 0) {
    int t = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += u;
    }
    wsum[lane] = t;
  }
  __syncthreads();
  const int incl = v + (w > 0 ? wsum[w - 1] : 0);
  if ((int)threadIdx.x < n_slots) pre[threadIdx.x + 1] = incl;
  if (threadIdx.x == 0) pre[0] = 0;
  __syncthreads();
  return *flag_s ? -1 : pre[n_slots];
}
// entry e of the row's concatenated segments
__device__ __forceinline__ unsigned long long segment_entry(const unsigned long long* __restrict__ buf_row, const int* pre,
                                                            int n_slots, int seg_cap, int e) {
  int lo = 0, hi = n_slots;                       // largest s with pre[s] <= e
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (pre[mid] <= e) lo = mid; else hi = mid;
  }
  return buf_row[(size_t)lo * seg_cap + (e - pre[lo])];
}

// Between the two scan phases: the K-th largest of the survivors the first eighth of the tiles produced is a
// (much) tighter threshold that K candidates are still guaranteed to reach.  Rows that overflowed keep theirs.
__global__ void __launch_bounds__(SC_SEL_THREADS)
topk_tau_refine_kernel(const unsigned long long* __restrict__ buf, const int* __restrict__ cnt, int n_slots, int seg_cap,
                       int K, float* __restrict__ tau) {
  __shared__ unsigned hist[257];
  __shared__ int bc[2];
  __shared__ int pre[SC_MAX_SLOTS + 1];
  __shared__ int over;
  const int row = blockIdx.x;
  const int n = segment_prefix(cnt + (size_t)row * n_slots, n_slots, seg_cap, pre, &over);
  if (n < K || n > SC_SEL_NPT * SC_SEL_THREADS) return;           // overflow (-1) included: the row keeps its threshold
  const unsigned long long* src = buf + (size_t)row * n_slots * seg_cap;
  unsigned key[SC_SEL_NPT];                       // the score half of the 64-bit keys orders them well enough here
#pragma unroll
  for (int i = 0; i < SC_SEL_NPT; ++i) {
    const int e = threadIdx.x + i * SC_SEL_THREADS;
    key[i] = e < n ? (unsigned)(segment_entry(src, pre, n_slots, seg_cap, e) >> 32) : 0u;
  }
  const unsigned kk = block_kth_largest<unsigned, SC_SEL_NPT>(key, n, K, hist, bc);
  if (threadIdx.x == 0) tau[row] = fmaxf(tau[row], just_below(key32_score(kk)));
}

// One CTA per query row: the row's survivors -> the best kp as (score desc, index asc), raw candidate indices:
// radix select of the kp-th largest key, compaction of the keys at or above it, sort of those kp.
// A row with more than cap survivors (or fewer than kp) raises the flag: the list-keeping path redoes the batch.
__global__ void __launch_bounds__(SC_SEL_THREADS)
topk_pool_select_kernel(const unsigned long long* __restrict__ buf, const int* __restrict__ cnt, int n_slots, int seg_cap,
                        int kp, int P, float* __restrict__ pool_s, int64_t* __restrict__ pool_i, int* __restrict__ flag) {
  extern __shared__ __align__(16) uint8_t sel_smem[];
  __shared__ unsigned hist[257];
  __shared__ int bc[2];
  __shared__ int n_out;
  __shared__ int pre[SC_MAX_SLOTS + 1];
  __shared__ int over;
  unsigned long long* list = reinterpret_cast<unsigned long long*>(sel_smem);      // [P], P = 2^m >= kp
  const int row = blockIdx.x;
  const int n_raw = segment_prefix(cnt + (size_t)row * n_slots, n_slots, seg_cap, pre, &over);
  if (n_raw < kp || n_raw > SC_SEL_NPT * SC_SEL_THREADS) {        // a segment overflowed (-1), or too few / too many survivors
    if (threadIdx.x == 0) atomicExch(flag, 1);
    return;
  }
  const unsigned long long* src = buf + (size_t)row * n_slots * seg_cap;
  unsigned long long key[SC_SEL_NPT];
#pragma unroll
  for (int i = 0; i < SC_SEL_NPT; ++i) {
    const int e = threadIdx.x + i * SC_SEL_THREADS;
    key[i] = e < n_raw ? segment_entry(src, pre, n_slots, seg_cap, e) : 0ull;
  }
  if (threadIdx.x == 0) n_out = 0;
  for (int i = threadIdx.x; i < P; i += SC_SEL_THREADS) list[i] = 0ull;
  const unsigned long long kth = block_kth_largest<unsigned long long, SC_SEL_NPT>(key, n_raw, kp, hist, bc);
#pragma unroll
  for (int i = 0; i < SC_SEL_NPT; ++i)
    if (key[i] >= kth && key[i] != 0ull) list[atomicAdd(&n_out, 1)] = key[i];      // keys are distinct: exactly kp of them
  __syncthreads();
  block_bitonic_desc(list, P);
  for (int t = threadIdx.x; t < kp; t += SC_SEL_THREADS) {
    const unsigned long long v = list[t];
    pool_s[(size_t)row * kp + t] = v ? topk_key_score(v) : -INFINITY;
    pool_i[(size_t)row * kp + t] = v ? (int64_t)topk_key_index(v) : LLONG_MAX;
  }
}

// ---- host ----------------------------------------------------------------------------------------------------
constexpr int SC_CAP = 16384;          // survivors per row the select kernel holds in registers (16 keys per thread)
constexpr int SC_EXPECT = 3712;        // planned survivors per row (cap / 4.4)

static int splits_for(int64_t q_groups, int64_t units, int64_t min_units) {
  const int64_t waves = std::min<int64_t>(16, std::max<int64_t>(1, q_groups / 4));
  int64_t s = ceil_div((int64_t)num_sms() * waves, q_groups);
  s = std::min<int64_t>(s, std::max<int64_t>(1, units / min_units));
  return (int)std::max<int64_t>(1, s);
}

// TT_TOPK_SCAN=0 / tt_debug_topk_scan_mode(0): always the list-keeping kernel; mode 2: survivor buffers of 64 entries,
// so every row overflows and the device-side fallback runs (tests)
static int g_scan_mode = -1;
static int scan_mode() {
  if (g_scan_mode < 0) {
    const char* e = getenv("TT_TOPK_SCAN");
    g_scan_mode = (e && e[0] == '0') ? 0 : 1;
  }
  return g_scan_mode;
}

ScanPlan tc_topk_scan_plan(int64_t nq, int64_t nc, int64_t d, int kp) {
  ScanPlan p{};
  if (scan_mode() == 0) return p;
  if (d % 64 != 0 || d < 64 || d > 128 || nc >= INT_MAX - 256 || nq >= (1 << 30)) return p;
  const int64_t total_tiles = ceil_div(nc, SC_BN);
  const int64_t groups_all = total_tiles * 4;
  // G group maxima -> about kp * groups_all / G survivors per row; at least 4 kp groups so that the K-th largest is
  // an interior order statistic
  int64_t G = std::max<int64_t>(4 * (int64_t)kp, ceil_div((int64_t)kp * groups_all, SC_EXPECT));
  int64_t n_samp = ceil_div(G, 4);
  if (n_samp * 2 > total_tiles || 4 * n_samp > 32768) return p;       // the sample would be most of the scan / beyond the sort
  p.use = true;
  p.nqt = nq > SC_BM ? 2 : 1;
  p.q_groups = (int)ceil_div(nq, p.nqt * SC_BM);
  p.nq_pad = (int)round_up(nq, 256);
  p.total_tiles = (int)total_tiles;
  p.n_samp = (int)n_samp;
  p.G = (int)(4 * n_samp);
  p.P = 64;
  while (p.P < p.G) p.P <<= 1;
  p.kp = kp;
  const int ss = splits_for(p.q_groups, n_samp, 8);
  p.samp_tps = (int)ceil_div(n_samp, ss);
  p.samp_splits = (int)ceil_div(n_samp, p.samp_tps);
  // more than one query tile: the scan is bound by the tensor pipe and the selection warps, not by HBM, and every
  // survivor costs issue slots -- scan an eighth of the tiles, tighten the threshold on what they produced, scan the rest
  p.two_phase = p.nqt == 2 && total_tiles >= 64;
  const int64_t n_first = ceil_div(total_tiles, 8);
  const int64_t n_main = p.two_phase ? total_tiles - n_first : total_tiles;
  const int cs = splits_for(p.q_groups, n_main, 32);
  p.scan_tps = (int)ceil_div(n_main, cs);
  p.scan_splits = (int)ceil_div(n_main, p.scan_tps);
  const int fs = splits_for(p.q_groups, n_first, 16);
  p.first_tps = (int)ceil_div(n_first, fs);
  p.first_splits = (int)ceil_div(n_first, p.first_tps);
  // one survivor segment per (row, candidate split of either phase); mode 2 (tests): 2-entry segments, so every row overflows
  p.n_slots = p.scan_splits + (p.two_phase ? p.first_splits : 0);
  p.seg_cap = scan_mode() == 2 ? 2 : SC_CAP / p.n_slots;
  p.cap = p.n_slots * p.seg_cap;
  if (p.n_slots > SC_MAX_SLOTS || p.seg_cap < 2) { p.use = false; return p; }
  int64_t off = 0;
  p.off_samp = off; off += round_up((int64_t)p.G * nq * 4, 256);
  p.off_tau = off;  off += round_up((int64_t)p.nq_pad * 4, 256);
  p.off_cnt = off;  off += round_up(nq * (int64_t)p.n_slots * 4, 256);
  p.off_flag = off; off += 256;
  p.off_buf = off;  off += round_up(nq * (int64_t)p.cap * 8, 256);
  p.bytes = off;
  return p;
}

template <int NQT, bool SAMPLE>
static int launch_scan(const CUtensorMap& tmQ, const CUtensorMap& tmC, ScanArgs a, dim3 grid, cudaStream_t st) {
  const ScanLayout L = scan_layout(a.d, NQT, SAMPLE);
  if (L.stages < 2) return set_error(TT_ERR_UNSUPPORTED, "tt_topk_bruteforce(bf16): d=%d does not fit the scan pipeline", a.d);
  a.stages = L.stages;
  TT_CUDA_OK(cudaFuncSetAttribute(topk_scan_kernel<NQT, SAMPLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
  TT_PROF(SAMPLE ? "topk_scan_kernel(sample)" : "topk_scan_kernel", st);
  topk_scan_kernel<NQT, SAMPLE><<<grid, 32 * (1 + NQT) + NQT * 128, L.total, st>>>(tmQ, tmC, a);
  TT_LAUNCH_OK("topk_scan_kernel");
  return TT_OK;
}

// Steps 1-4 into pool_s / pool_i ([nq, kp], raw candidate indices).  *flag (device, inside ws) is 1 afterwards iff the
// caller's list-keeping path has to redo the batch.
int tc_topk_scan(const ScanPlan& p, const void* queries, const void* candidates, int64_t nq, int64_t nc, int64_t d,
                 float* pool_s, int64_t* pool_i, void* ws, int** flag_out, cudaStream_t st) {
  char* base = (char*)ws;
  float* samp = (float*)(base + p.off_samp);
  float* tau = (float*)(base + p.off_tau);
  int* cnt = (int*)(base + p.off_cnt);
  int* flag = (int*)(base + p.off_flag);
  unsigned long long* buf = (unsigned long long*)(base + p.off_buf);
  *flag_out = flag;
  CUtensorMap tmQ, tmC;
  int rc = make_tmap_bf16_2d(&tmQ, queries, (uint64_t)d, (uint64_t)nq, (uint64_t)d * 2, 64, SC_BM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmC, candidates, (uint64_t)d, (uint64_t)nc, (uint64_t)d * 2, 64, SC_BN);
  if (rc) return rc;
  ScanArgs a{};
  a.nq = (int)nq; a.nc = (int)nc; a.d = (int)d;
  a.total_tiles = p.total_tiles; a.n_samp = p.n_samp; a.nq_pad = p.nq_pad;
  a.n_slots = p.n_slots; a.seg_cap = p.seg_cap; a.slot0 = 0;
  a.tau = tau; a.cnt = cnt; a.buf = buf; a.samp = samp;
  a.trace = nullptr;
  a.tiles_per_split = p.samp_tps;
  const dim3 gs((unsigned)p.q_groups, (unsigned)p.samp_splits);
  rc = p.nqt == 2 ? launch_scan<2, true>(tmQ, tmC, a, gs, st) : launch_scan<1, true>(tmQ, tmC, a, gs, st);
  if (rc) return rc;
  TT_PROF("topk_tau_kernel", st);
  topk_tau_kernel<<<(unsigned)nq, SC_SEL_THREADS, 0, st>>>(samp, p.G, p.kp, tau, cnt, p.n_slots, flag);
  TT_LAUNCH_OK("topk_tau_kernel");
  if (!p.two_phase) {
    a.phase = 0; a.tiles_per_split = p.scan_tps;
    const dim3 gc((unsigned)p.q_groups, (unsigned)p.scan_splits);
    rc = p.nqt == 2 ? launch_scan<2, false>(tmQ, tmC, a, gc, st) : launch_scan<1, false>(tmQ, tmC, a, gc, st);
    if (rc) return rc;
  } else {
    a.phase = 1; a.tiles_per_split = p.first_tps;
    rc = launch_scan<2, false>(tmQ, tmC, a, dim3((unsigned)p.q_groups, (unsigned)p.first_splits), st);
    if (rc) return rc;
    TT_PROF("topk_tau_refine_kernel", st);
    topk_tau_refine_kernel<<<(unsigned)nq, SC_SEL_THREADS, 0, st>>>(buf, cnt, p.n_slots, p.seg_cap, p.kp, tau);
    TT_LAUNCH_OK("topk_tau_refine_kernel");
    a.phase = 2; a.tiles_per_split = p.scan_tps; a.slot0 = p.first_splits;
    a.trace = g_scan_trace;
    rc = launch_scan<2, false>(tmQ, tmC, a, dim3((unsigned)p.q_groups, (unsigned)p.scan_splits), st);
    if (rc) return rc;
  }
  int P = 32;
  while (P < p.kp) P <<= 1;
  TT_PROF("topk_pool_select_kernel", st);
  topk_pool_select_kernel<<<(unsigned)nq, SC_SEL_THREADS, P * 8, st>>>(buf, cnt, p.n_slots, p.seg_cap, p.kp, P, pool_s, pool_i, flag);
  TT_LAUNCH_OK("topk_pool_select_kernel");
  return TT_OK;
}

}  // namespace tt

extern "C" int tt_debug_topk_scan_trace(long long* device_buf) {
  tt::g_scan_trace = device_buf;
  return TT_OK;
}

extern "C" int tt_debug_topk_scan_mode(int32_t mode) {
  const int prev = tt::scan_mode();
  if (mode >= 0 && mode <= 2) tt::g_scan_mode = mode;
  return prev;
}
