// K3/K4 (bf16 precision) -- tfrs.tasks.Retrieval in-batch softmax cross-entropy on tcgen05
// tensor cores, flash-attention style: the [nq, nc] logits live only in TMEM.
//
// forward  (retrieval_fwd_tc_kernel): a CTA keeps a 128-query tile resident in shared memory
//   and streams 128-candidate tiles through a TMA ring.  S = Q C^T is accumulated in TMEM
//   (double-buffered, 2 x 128 columns); two softmax warpgroups alternate tiles, each thread
//   owning one query row (tcgen05.ld 32x32b: no cross-thread reductions), keeping a running
//   (max, sum-exp) in the log2 domain.  Candidate ranges are split over CTAs to fill the SMs;
//   partial (max, sum) pairs are merged by a finalize kernel into row_lse and the SUM loss.
//
// backward (retrieval_bwd_tc_kernel<TRANSPOSED>): the same streaming structure, plus a second
//   GEMM per tile.  The softmax warps turn S into dS = softmax - eye (bf16) and store it to
//   shared memory in the 128B-swizzled K-major layout; the MMA thread then accumulates
//   dX[128, d] += dS[128, BN] * Y-tile into a TMEM accumulator, reading the SAME shared-memory Y
//   tile as an MN-major B operand (no transposed copy is loaded or kept in HBM).  TRANSPOSED = false keeps
//   query rows stationary (dQ); TRANSPOSED = true keeps candidate rows stationary and streams
//   queries (dC).  S is recomputed in each pass (5 GEMM units for 3 algorithmic ones).
//
// Upstream order of the logit transforms is kept: / temperature, - log q, accidental-hit
// mask; sample weights multiply dS.  All exponentials are ex2.approx on log2-scaled logits.
#include "tc_common.cuh"

namespace tt {

int launch_loss_reduce(const float* lse, const float* pos, const float* w, int64_t n, float* loss, cudaStream_t st);

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int RT_BM = 128;           // stationary rows per CTA
constexpr int RT_THREADS = 320;      // warp 0 TMA, warp 1 MMA, warps 2-5 / 6-9 softmax warpgroups

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct RetrievalTcArgs {
  int nq, nc, d;
  float k2;                       // inv_temperature * log2(e)
  float out_scale;                // inv_temperature * grad_scale (backward)
  long long label_offset;
  const float* w;                 // sample weights [nq] or null
  const float* logq;              // log(clip(p)) per candidate [nc] or null
  const long long* cand_ids;      // [nc] or null (accidental-hit removal)
  const float* lse;               // [nq] natural-log lse (backward)
  float2* partial_ml;             // forward: [splits][nq] (max2, sum)
  float* row_pos;                 // forward: [nq]
  float* partial_out;             // backward: [splits][nX][d]
  int tiles_per_split;            // streamed tiles per CTA (blockIdx.y)
  int stages;                     // depth of the streamed-tile ring (<= 4)
};

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
template <int BN>
struct FwdSmem {
  __host__ __device__ static constexpr int q_bytes(int d) { return RT_BM * d * 2; }
  __host__ __device__ static constexpr int y_bytes(int d) { return BN * d * 2; }
  __host__ __device__ static constexpr int stages(int d) {
    return (227 * 1024 - 1024 - 8192 - q_bytes(d)) / y_bytes(d) > 4 ? 4 : (227 * 1024 - 1024 - 8192 - q_bytes(d)) / y_bytes(d);
  }
  __host__ __device__ static constexpr int total(int d) { return q_bytes(d) + stages(d) * y_bytes(d) + 8192 + 1024; }
};

template <int BN, bool EXTRAS>
__global__ void __launch_bounds__(RT_THREADS, 1)
retrieval_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmC,
                        const RetrievalTcArgs a) {
  using L = FwdSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int d = a.d, nkb = d / 64;
  const int STAGES = a.stages;
  uint8_t* sQ = smem;
  uint8_t* sY = sQ + L::q_bytes(d);
  uint8_t* tail = sY + STAGES * L::y_bytes(d);
  uint64_t* q_full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* full = q_full + 1;
  uint64_t* empty = full + 4;
  uint64_t* s_full = empty + 4;
  uint64_t* s_empty = s_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + 2);
  float2* wg_ml = reinterpret_cast<float2*>(tail + 256);          // [128] partials of warpgroup 1
  float* col_logq2 = reinterpret_cast<float*>(tail + 256 + 1024); // [2][BN]   (EXTRAS)
  long long* col_id = reinterpret_cast<long long*>(tail + 256 + 1024 + 1024);   // [2][BN] needs 2 KB

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * RT_BM;
  const int tile_begin = blockIdx.y * a.tiles_per_split;
  const int total_tiles = (a.nc + BN - 1) / BN;
  const int T = max(0, min(a.tiles_per_split, total_tiles - tile_begin));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmC);
    mbar_init(q_full, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&s_full[b], 1); mbar_init(&s_empty[b], 128); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, L::q_bytes(d));
      for (int kb = 0; kb < nkb; ++kb) tma_load_2d(sQ + kb * RT_BM * 128, &tmQ, q_full, kb * 64, q0);
      for (int t = 0; t < T; ++t) {
        const int s = t % STAGES;
        mbar_wait(&empty[s], ((t / STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[s], L::y_bytes(d));
        for (int kb = 0; kb < nkb; ++kb)
          tma_load_2d(sY + s * L::y_bytes(d) + kb * BN * 128, &tmC, &full[s], kb * 64, (tile_begin + t) * BN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(RT_BM, BN);
      mbar_wait(q_full, 0);
      for (int t = 0; t < T; ++t) {
        const int s = t % STAGES, b = t & 1;
        mbar_wait(&full[s], (t / STAGES) & 1);
        mbar_wait(&s_empty[b], ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          const uint64_t da = umma_desc_k_sw128(smem_u32(sQ + kb * RT_BM * 128));
          const uint64_t db = umma_desc_k_sw128(smem_u32(sY + s * L::y_bytes(d) + kb * BN * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + b * BN, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
        }
        umma_commit(&s_full[b]);
        umma_commit(&empty[s]);
      }
    }
  } else {
    const int g = (warp - 2) >> 2;                 // softmax warpgroup 0 / 1
    const int qd = warp & 3;                       // TMEM lane quarter
    const int r = qd * 32 + lane;                  // row inside the tile
    const int wg_tid = ((warp - 2) & 3) * 32 + lane;
    const long long qi = (long long)q0 + r;
    const long long label = a.label_offset + qi;
    long long pos_id = -1;
    if (EXTRAS && a.cand_ids && qi < a.nq) pos_id = a.cand_ids[label];
    float m2 = -INFINITY, l = 0.f;
    for (int t = g; t < T; t += 2) {
      const int b = t & 1;                         // == g
      const long long c_tile = (long long)(tile_begin + t) * BN;
      if (EXTRAS) {
        // this warpgroup's 128 threads stage the per-column terms of the tile
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
        if (wg_tid < BN) {
          const long long ci = c_tile + wg_tid;
          col_logq2[b * BN + wg_tid] = (a.logq && ci < a.nc) ? a.logq[ci] * kLog2e : 0.f;
          col_id[b * BN + wg_tid] = (a.cand_ids && ci < a.nc) ? a.cand_ids[ci] : -2;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
      }
      mbar_wait(&s_full[b], (t >> 1) & 1);
      tc_fence_after();
      // whole row of the tile into registers, then hand the TMEM buffer back at once so the
      // next tile's MMA overlaps the exponentials
      uint32_t rr[BN];
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + b * BN + c * 32, rr + c * 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_empty[b]);
      const bool edge = c_tile + BN > a.nc;
      const bool has_label = label >= c_tile && label < c_tile + BN;
      if (!EXTRAS && !edge && !has_label) {
        // fast path: max on the raw accumulators (k2 > 0), then one FFMA + ex2 + add per logit
        float cmax = fmax3(__uint_as_float(rr[0]), __uint_as_float(rr[1]), __uint_as_float(rr[2]));
#pragma unroll
        for (int j = 3; j + 1 < BN; j += 2) cmax = fmax3(cmax, __uint_as_float(rr[j]), __uint_as_float(rr[j + 1]));
        cmax = fmaxf(cmax, __uint_as_float(rr[BN - 1])) * a.k2;
        if (cmax > m2) { l *= ex2_approx(m2 - cmax); m2 = cmax; }
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        const float nm = -m2;
#pragma unroll
        for (int j = 0; j < BN; j += 4) {
          acc0 += ex2_approx(fmaf(__uint_as_float(rr[j]), a.k2, nm));
          acc1 += ex2_approx(fmaf(__uint_as_float(rr[j + 1]), a.k2, nm));
          acc2 += ex2_approx(fmaf(__uint_as_float(rr[j + 2]), a.k2, nm));
          acc3 += ex2_approx(fmaf(__uint_as_float(rr[j + 3]), a.k2, nm));
        }
        l += (acc0 + acc1) + (acc2 + acc3);
      } else {
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
          float x[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(rr[c0 + j]) * a.k2;
          if (EXTRAS) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              x[j] -= col_logq2[b * BN + c0 + j];
              const long long ci = c_tile + c0 + j;
              if (col_id[b * BN + c0 + j] == pos_id && ci != label) x[j] += TT_MIN_FLOAT;
            }
          }
          if (has_label && label >= c_tile + c0 && label < c_tile + c0 + 32) {
            const int jj = (int)(label - c_tile - c0);
            float p = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j == jj) p = x[j];
            if (qi < a.nq) a.row_pos[qi] = p * kLn2;
          }
          if (edge) {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (c_tile + c0 + j >= a.nc) x[j] = -INFINITY;
          }
          float cmax = x[0];
#pragma unroll
          for (int j = 1; j < 32; ++j) cmax = fmaxf(cmax, x[j]);
          if (cmax > m2) { l *= ex2_approx(m2 - cmax); m2 = cmax; }
          if (m2 > -INFINITY) {
            float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; j += 2) { acc0 += ex2_approx(x[j] - m2); acc1 += ex2_approx(x[j + 1] - m2); }
            l += acc0 + acc1;
          }
        }
      }
    }
    // merge the two warpgroups' partials, write one (max, sum) per row and split
    if (g == 1) wg_ml[r] = make_float2(m2, l);
    asm volatile("bar.sync 3, 256;" ::: "memory");
    if (g == 0 && qi < a.nq) {
      const float2 o = wg_ml[r];
      const float mn = fmaxf(m2, o.x);
      float ln = 0.f;
      if (mn > -INFINITY) ln = l * ex2_approx(m2 - mn) + o.y * ex2_approx(o.x - mn);
      a.partial_ml[(size_t)blockIdx.y * a.nq + qi] = make_float2(mn, ln);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * BN);
}

// lse_i = ln2 * (M + log2(sum_s l_s 2^(m_s - M))) over the candidate splits
__global__ void __launch_bounds__(256)
retrieval_fwd_finalize_kernel(const float2* __restrict__ partial, int splits, int nq, float* __restrict__ row_lse) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  float M = -INFINITY;
  for (int s = 0; s < splits; ++s) M = fmaxf(M, partial[(size_t)s * nq + i].x);
  float L = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float2 p = partial[(size_t)s * nq + i];
    if (p.x > -INFINITY) L += p.y * exp2f(p.x - M);
  }
  row_lse[i] = (M + log2f(L)) * kLn2;
}

// ---------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------
struct BwdLayout {
  int x_bytes, y_bytes, ds_bytes, stage_bytes, stages, total;
};
// tail = barriers (256 B) + per-column lse/weight vectors (2 KB) + per-column ids (2 KB)
__host__ __device__ inline int bwd_tail_bytes(bool transposed, bool extras) {
  return 256 + ((transposed || extras) ? 2048 : 0) + (extras ? 2048 : 0);
}
__host__ __device__ inline BwdLayout bwd_layout(int d, int BN, int tail_bytes) {
  BwdLayout L;
  L.x_bytes = RT_BM * d * 2;
  L.y_bytes = BN * d * 2;
  L.ds_bytes = RT_BM * BN * 2;
  L.stage_bytes = L.y_bytes;
  const int budget = 227 * 1024 - tail_bytes - L.x_bytes - 2 * L.ds_bytes;
  L.stages = budget / L.stage_bytes;
  if (L.stages > 4) L.stages = 4;
  L.total = L.x_bytes + 2 * L.ds_bytes + L.stages * L.stage_bytes + tail_bytes;
  return L;
}

template <int BN, bool TRANSPOSED, bool EXTRAS>
__global__ void __launch_bounds__(RT_THREADS, 1)
retrieval_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                        const RetrievalTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;      // no slack: the swizzled tiles need the 1024-byte alignment the declaration asks for
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  const int d = a.d, nkb = d / 64;
  const BwdLayout L = bwd_layout(d, BN, bwd_tail_bytes(TRANSPOSED, EXTRAS));
  const int STAGES = L.stages;
  uint8_t* sX = smem;
  uint8_t* sDS = sX + L.x_bytes;                      // [2][BN/64][128 x 64] bf16, SW128
  uint8_t* sY = sDS + 2 * L.ds_bytes;                 // per stage: Y tile [d/64][BN x 64], SW128
  uint8_t* tail = sY + STAGES * L.stage_bytes;
  uint64_t* x_full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* full = x_full + 1;
  uint64_t* empty = full + 4;
  uint64_t* s_full = empty + 4;
  uint64_t* s_empty = s_full + 2;
  uint64_t* ds_full = s_empty + 2;
  uint64_t* ds_empty = ds_full + 2;
  uint64_t* acc_full = ds_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  float* col_a = reinterpret_cast<float*>(tail + 256);              // [2][BN] lse2 (TRANSPOSED) or logq2
  float* col_w = reinterpret_cast<float*>(tail + 256 + 1024);       // [2][BN] weights (TRANSPOSED)
  long long* col_id = reinterpret_cast<long long*>(tail + 256 + 2048);   // [2][BN] cand ids / positive ids (EXTRAS)

  const int nX = TRANSPOSED ? a.nc : a.nq;
  const int nY = TRANSPOSED ? a.nq : a.nc;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int x0 = blockIdx.x * RT_BM;
  const int tile_begin = blockIdx.y * a.tiles_per_split;
  const int total_tiles = (nY + BN - 1) / BN;
  const int T = max(0, min(a.tiles_per_split, total_tiles - tile_begin));
  const uint32_t ACC_COL = 2 * BN;                    // TMEM: [S0 | S1 | acc(d)]

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmY);
    mbar_init(x_full, 1);
    for (int s = 0; s < 4; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1); mbar_init(&s_empty[b], 128);
      mbar_init(&ds_full[b], 128); mbar_init(&ds_empty[b], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(x_full, L.x_bytes);
      for (int kb = 0; kb < nkb; ++kb) tma_load_2d(sX + kb * RT_BM * 128, &tmX, x_full, kb * 64, x0);
      for (int t = 0; t < T; ++t) {
        const int s = t % STAGES;
        mbar_wait(&empty[s], ((t / STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[s], L.stage_bytes);
        uint8_t* base = sY + s * L.stage_bytes;
        const int y0 = (tile_begin + t) * BN;
        for (int kb = 0; kb < nkb; ++kb) tma_load_2d(base + kb * BN * 128, &tmY, &full[s], kb * 64, y0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && T > 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(RT_BM, BN);
      const uint32_t idesc2 = umma_idesc_bf16(RT_BM, d, 0, 1);     // B = Y tile read MN-major (N = d contiguous)
      mbar_wait(x_full, 0);
      auto issue_mma1 = [&](int t) {
        const int s = t % STAGES, b = t & 1;
        mbar_wait(&full[s], (t / STAGES) & 1);
        mbar_wait(&s_empty[b], ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          const uint64_t da = umma_desc_k_sw128(smem_u32(sX + kb * RT_BM * 128));
          const uint64_t db = umma_desc_k_sw128(smem_u32(sY + s * L.stage_bytes + kb * BN * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + b * BN, da + 2 * k, db + 2 * k, idesc1, (kb | k) != 0);
        }
        umma_commit(&s_full[b]);
      };
      issue_mma1(0);
      for (int t = 0; t < T; ++t) {
        if (t + 1 < T) issue_mma1(t + 1);
        const int s = t % STAGES, b = t & 1;
        mbar_wait(&ds_full[b], (t >> 1) & 1);
        tc_fence_after();
        // K = the BN streamed rows: 16 rows (2048 B) per MMA; the d/64 column chunks of the Y tile
        // are BN*128 bytes apart (LBO)
        const uint64_t db0 = umma_desc_mn_sw128(smem_u32(sY + s * L.stage_bytes), BN * 128);
        for (int jb = 0; jb < BN / 64; ++jb) {
          const uint64_t da = umma_desc_k_sw128(smem_u32(sDS + b * L.ds_bytes + jb * RT_BM * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tmem_base + ACC_COL, da + 2 * k, db0 + 128 * (jb * 4 + k), idesc2, (t | jb | k) != 0);
        }
        umma_commit(&ds_empty[b]);
        umma_commit(&empty[s]);
      }
      umma_commit(acc_full);
    }
  } else {
    const int g = (warp - 2) >> 2;
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const int wg_tid = ((warp - 2) & 3) * 32 + lane;
    const long long xi = (long long)x0 + r;                        // query (dQ) or candidate (dC) index
    // per-row constants
    float row_lse2 = 0.f, row_scale = 0.f, row_logq2 = 0.f;
    long long row_id = -1;                                         // dQ: positive id of the query; dC: id of the candidate
    if (xi < nX) {
      if (!TRANSPOSED) {
        row_lse2 = a.lse[xi] * kLog2e;
        row_scale = (a.w ? a.w[xi] : 1.f) * a.out_scale;
        if (EXTRAS && a.cand_ids) row_id = a.cand_ids[a.label_offset + xi];
      } else {
        row_scale = a.out_scale;
        if (EXTRAS && a.logq) row_logq2 = a.logq[xi] * kLog2e;
        if (EXTRAS && a.cand_ids) row_id = a.cand_ids[xi];
      }
    }
    const bool col_weighted = TRANSPOSED && a.w != nullptr;
    for (int t = g; t < T; t += 2) {
      const int b = t & 1;
      const long long y_tile = (long long)(tile_begin + t) * BN;
      if (TRANSPOSED || EXTRAS) {
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
        if (wg_tid < BN) {
          const long long yi = y_tile + wg_tid;
          if (TRANSPOSED) {
            col_a[b * BN + wg_tid] = yi < a.nq ? a.lse[yi] * kLog2e : 0.f;
            col_w[b * BN + wg_tid] = (a.w && yi < a.nq) ? a.w[yi] : 1.f;
            if (EXTRAS) col_id[b * BN + wg_tid] = (a.cand_ids && yi < a.nq) ? a.cand_ids[a.label_offset + yi] : -2;
          } else {
            col_a[b * BN + wg_tid] = (a.logq && yi < a.nc) ? a.logq[yi] * kLog2e : 0.f;
            col_id[b * BN + wg_tid] = (a.cand_ids && yi < a.nc) ? a.cand_ids[yi] : -2;
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
      }
      mbar_wait(&s_full[b], (t >> 1) & 1);
      tc_fence_after();
      uint32_t rr[BN];
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + b * BN + c * 32, rr + c * 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_empty[b]);                      // S buffer free: the MMA of tile t+2 may start
      mbar_wait(&ds_empty[b], ((t >> 1) & 1) ^ 1);   // dS buffer free (MMA2 of tile t-2 done)
      // label column (dQ) / label row (dC) intersects this tile?
      // dQ: element (xi, y) is the positive when y == label_offset + xi  <=>  xi == y - label_offset
      // dC: element (xi, y) is the positive when xi == label_offset + y
      const long long lab_lo = TRANSPOSED ? y_tile + a.label_offset : y_tile - a.label_offset;
      const bool diag = xi >= lab_lo && xi < lab_lo + BN;
      uint8_t* ds_base = sDS + b * L.ds_bytes;
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float p[32];
        if (!TRANSPOSED) {
#pragma unroll
          for (int j = 0; j < 32; ++j) p[j] = fmaf(__uint_as_float(rr[c0 + j]), a.k2, -row_lse2);
          if (EXTRAS) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              p[j] -= col_a[b * BN + c0 + j];
              if (col_id[b * BN + c0 + j] == row_id && (y_tile + c0 + j) != a.label_offset + xi) p[j] = -INFINITY;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 l4 = *reinterpret_cast<const float4*>(col_a + b * BN + c0 + j);
            p[j] = fmaf(__uint_as_float(rr[c0 + j]), a.k2, -l4.x - row_logq2);
            p[j + 1] = fmaf(__uint_as_float(rr[c0 + j + 1]), a.k2, -l4.y - row_logq2);
            p[j + 2] = fmaf(__uint_as_float(rr[c0 + j + 2]), a.k2, -l4.z - row_logq2);
            p[j + 3] = fmaf(__uint_as_float(rr[c0 + j + 3]), a.k2, -l4.w - row_logq2);
          }
          if (EXTRAS) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col_id[b * BN + c0 + j] == row_id && xi != a.label_offset + y_tile + c0 + j) p[j] = -INFINITY;
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) p[j] = ex2_approx(p[j]);
        if (diag) {
          const long long jj = xi - lab_lo - c0;
#pragma unroll
          for (int j = 0; j < 32; ++j) if (j == jj) p[j] -= 1.f;
        }
        if (col_weighted) {
#pragma unroll
          for (int j = 0; j < 32; ++j) p[j] *= col_w[b * BN + c0 + j];
        }
        // bf16 pack + swizzled store: 16-byte chunk (c0/8 + g8) of row r in sub-tile (c0 / 64)
        uint8_t* sub = ds_base + (c0 >> 6) * (RT_BM * 128);
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const uint4 v = make_uint4(pack_bf16x2(p[g8 * 8], p[g8 * 8 + 1]), pack_bf16x2(p[g8 * 8 + 2], p[g8 * 8 + 3]),
                                     pack_bf16x2(p[g8 * 8 + 4], p[g8 * 8 + 5]), pack_bf16x2(p[g8 * 8 + 6], p[g8 * 8 + 7]));
          *reinterpret_cast<uint4*>(sub + sw128_offset(r, ((c0 & 63) >> 3) + g8)) = v;
        }
      }
      fence_proxy_async();
      mbar_arrive(&ds_full[b]);
    }
    // epilogue: accumulator [128 x d] -> fp32 partial; warpgroup g takes column half g
    if (T > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
    }
    float* out = a.partial_out + ((size_t)blockIdx.y * nX + (size_t)xi) * d;
    const int half = d / 2;
#pragma unroll 1
    for (int c0 = g * half; c0 < (g + 1) * half; c0 += 32) {
      uint32_t rr[32];
      if (T > 0) {
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(qd * 32) << 16) + ACC_COL + c0, rr);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) rr[j] = 0u;
      }
      if (xi < nX) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(out + c0 + j) =
              make_float4(__uint_as_float(rr[j]) * row_scale, __uint_as_float(rr[j + 1]) * row_scale,
                          __uint_as_float(rr[j + 2]) * row_scale, __uint_as_float(rr[j + 3]) * row_scale);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// out = sum_s partial[s]; optional bf16 copy.  One warp per row.
__global__ void __launch_bounds__(256)
combine_partials_kernel(const float* __restrict__ partial, int splits, int64_t rows, int d, float* __restrict__ out_f32,
                        uint16_t* __restrict__ out_bf16) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  for (int c = lane * 4; c < d; c += 128) {
    float4 s = *reinterpret_cast<const float4*>(partial + r * d + c);
    for (int p = 1; p < splits; ++p) {
      const float4 v = *reinterpret_cast<const float4*>(partial + ((size_t)p * rows + r) * d + c);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + r * d + c) = s;
    if (out_bf16) *reinterpret_cast<uint2*>(out_bf16 + r * d + c) = make_uint2(pack_bf16x2(s.x, s.y), pack_bf16x2(s.z, s.w));
  }
}

// ---- host ----------------------------------------------------------------------------
// 128-candidate streamed tiles when they fit (d <= 128, no per-column id vectors), else 64
static int bn_for(int64_t d, bool extras) { return (d <= 128 && !extras) ? 128 : 64; }

// The streamed range is split over blockIdx.y so that x_tiles * splits ~ fills the SMs once.
// `splits` depends only on 128-row granules (so workspace sizes do not depend on BN).
static void split_plan(int64_t nX, int64_t nY, int BN, int* splits, int* tiles_per_split) {
  const int64_t x_tiles = ceil_div(nX, RT_BM), y128 = ceil_div(nY, 128);
  int64_t s = std::max<int64_t>(1, num_sms() / x_tiles);
  if (s > y128) s = y128;
  const int64_t per128 = ceil_div(y128, s);
  *tiles_per_split = (int)(per128 * (128 / BN));
  *splits = (int)ceil_div(y128, per128);
}

struct WsPlan { int sf, sq, sc; int64_t off_q, off_c, total; };
static WsPlan ws_plan(int64_t nq, int64_t nc, int64_t d) {
  WsPlan p;
  int tps;
  split_plan(nq, nc, 128, &p.sf, &tps);
  split_plan(nq, nc, 128, &p.sq, &tps);
  split_plan(nc, nq, 128, &p.sc, &tps);
  p.off_q = round_up((int64_t)p.sf * nq * 8, 256);
  p.off_c = p.off_q + round_up((int64_t)p.sq * nq * d * 4, 256);
  p.total = p.off_c + round_up((int64_t)p.sc * nc * d * 4, 256);
  return p;
}

int64_t tc_retrieval_workspace_bytes(int64_t nq, int64_t nc, int64_t d) { return ws_plan(nq, nc, d).total; }

static int check_tc_dims(const char* fn, int64_t nq, int64_t nc, int64_t d) {
  TT_REQUIRE(d % 64 == 0 && d >= 64 && d <= 256, "%s(bf16): d must be 64, 128, 192 or 256 (got %lld)", fn, (long long)d);
  TT_REQUIRE(nq < (1ll << 31) && nc < (1ll << 31), "%s(bf16): sizes exceed int32", fn);
  return TT_OK;
}

int tc_retrieval_fwd(const void* q, const void* c, int64_t nq, int64_t nc, int64_t d, float inv_temp,
                     int64_t label_offset, const float* w, const float* logq, const int64_t* cand_ids,
                     float* row_lse, float* row_pos, float* loss, void* ws, int64_t ws_bytes, cudaStream_t st) {
  int rc = check_tc_dims("tt_retrieval_loss_fwd", nq, nc, d);
  if (rc) return rc;
  if (!ws || ws_bytes < tc_retrieval_workspace_bytes(nq, nc, d))
    return set_error(TT_ERR_WORKSPACE, "tt_retrieval_loss_fwd(bf16): workspace too small");
  constexpr int BN = 128;
  int splits, tps;
  split_plan(nq, nc, BN, &splits, &tps);
  CUtensorMap tmQ, tmC;
  rc = make_tmap_bf16_2d(&tmQ, q, (uint64_t)d, (uint64_t)nq, (uint64_t)d * 2, 64, RT_BM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmC, c, (uint64_t)d, (uint64_t)nc, (uint64_t)d * 2, 64, BN);
  if (rc) return rc;
  RetrievalTcArgs a{};
  a.nq = (int)nq; a.nc = (int)nc; a.d = (int)d;
  a.k2 = inv_temp * kLog2e;
  a.label_offset = label_offset; a.w = w; a.logq = logq; a.cand_ids = (const long long*)cand_ids;
  a.partial_ml = (float2*)ws; a.row_pos = row_pos; a.tiles_per_split = tps;
  a.stages = FwdSmem<BN>::stages((int)d);
  TT_REQUIRE(a.stages >= 2, "tt_retrieval_loss_fwd(bf16): d=%lld does not fit the shared-memory pipeline", (long long)d);
  const int smem = FwdSmem<BN>::total((int)d);
  dim3 grid((unsigned)ceil_div(nq, RT_BM), (unsigned)splits);
  if (logq || cand_ids) {
    TT_CUDA_OK(cudaFuncSetAttribute(retrieval_fwd_tc_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    TT_PROF("retrieval_fwd_tc_kernel", st);
    retrieval_fwd_tc_kernel<BN, true><<<grid, RT_THREADS, smem, st>>>(tmQ, tmC, a);
  } else {
    TT_CUDA_OK(cudaFuncSetAttribute(retrieval_fwd_tc_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    TT_PROF("retrieval_fwd_tc_kernel", st);
    retrieval_fwd_tc_kernel<BN, false><<<grid, RT_THREADS, smem, st>>>(tmQ, tmC, a);
  }
  TT_LAUNCH_OK("retrieval_fwd_tc_kernel");
  TT_PROF("retrieval_fwd_finalize_kernel", st);
  retrieval_fwd_finalize_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, st>>>((const float2*)ws, splits, (int)nq, row_lse);
  TT_LAUNCH_OK("retrieval_fwd_finalize_kernel");
  return launch_loss_reduce(row_lse, row_pos, w, nq, loss, st);
}

template <int BN, bool TRANSPOSED>
static int launch_bwd(const void* x, const void* y, int64_t nX, int64_t nY, RetrievalTcArgs a,
                      float* partial, int* splits_out, cudaStream_t st) {
  const int d = a.d;
  int splits, tps;
  split_plan(nX, nY, BN, &splits, &tps);
  *splits_out = splits;
  CUtensorMap tmX, tmY;
  int rc = make_tmap_bf16_2d(&tmX, x, (uint64_t)d, (uint64_t)nX, (uint64_t)d * 2, 64, RT_BM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmY, y, (uint64_t)d, (uint64_t)nY, (uint64_t)d * 2, 64, BN);
  if (rc) return rc;
  a.partial_out = partial; a.tiles_per_split = tps;
  const bool extras = a.logq || a.cand_ids;
  const BwdLayout L = bwd_layout(d, BN, bwd_tail_bytes(TRANSPOSED, extras));
  TT_REQUIRE(L.stages >= 2, "tt_retrieval_loss_bwd(bf16): d=%d does not fit the shared-memory pipeline", d);
  dim3 grid((unsigned)ceil_div(nX, RT_BM), (unsigned)splits);
#define TT_BWD_LAUNCH(EX)                                                                                     \
  {                                                                                                           \
    TT_CUDA_OK(cudaFuncSetAttribute(retrieval_bwd_tc_kernel<BN, TRANSPOSED, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total)); \
    TT_PROF("retrieval_bwd_tc_kernel", st), retrieval_bwd_tc_kernel<BN, TRANSPOSED, EX><<<grid, RT_THREADS, L.total, st>>>(tmX, tmY, a);        \
  }
  if (extras) TT_BWD_LAUNCH(true) else TT_BWD_LAUNCH(false)
#undef TT_BWD_LAUNCH
  TT_LAUNCH_OK("retrieval_bwd_tc_kernel");
  return TT_OK;
}

int tc_retrieval_bwd(const void* q, const void* c, int64_t nq, int64_t nc,
                     int64_t d, float inv_temp, int64_t label_offset, const float* w, const float* logq,
                     const int64_t* cand_ids, const float* row_lse, float grad_scale, float* dq, float* dc,
                     uint16_t* dq_bf16, uint16_t* dc_bf16, void* ws, int64_t ws_bytes, cudaStream_t st) {
  int rc = check_tc_dims("tt_retrieval_loss_bwd", nq, nc, d);
  if (rc) return rc;
  if (!ws || ws_bytes < tc_retrieval_workspace_bytes(nq, nc, d))
    return set_error(TT_ERR_WORKSPACE, "tt_retrieval_loss_bwd(bf16): workspace too small");
  RetrievalTcArgs a{};
  a.nq = (int)nq; a.nc = (int)nc; a.d = (int)d;
  a.k2 = inv_temp * kLog2e; a.out_scale = inv_temp * grad_scale;
  a.label_offset = label_offset; a.w = w; a.logq = logq; a.cand_ids = (const long long*)cand_ids; a.lse = row_lse;
  const int BN = bn_for(d, logq || cand_ids);
  const WsPlan plan = ws_plan(nq, nc, d);
  float* part_q = (float*)((char*)ws + plan.off_q);
  float* part_c = (float*)((char*)ws + plan.off_c);
  int sq = 1, sc = 1;
  if (BN == 128) {
    rc = launch_bwd<128, false>(q, c, nq, nc, a, part_q, &sq, st);
    if (rc) return rc;
    rc = launch_bwd<128, true>(c, q, nc, nq, a, part_c, &sc, st);
  } else {
    rc = launch_bwd<64, false>(q, c, nq, nc, a, part_q, &sq, st);
    if (rc) return rc;
    rc = launch_bwd<64, true>(c, q, nc, nq, a, part_c, &sc, st);
  }
  if (rc) return rc;
  TT_PROF("combine_partials_kernel", st);
  combine_partials_kernel<<<(unsigned)ceil_div(nq, 8), 256, 0, st>>>(part_q, sq, nq, (int)d, dq, dq_bf16);
  TT_LAUNCH_OK("combine_partials_kernel");
  TT_PROF("combine_partials_kernel", st);
  combine_partials_kernel<<<(unsigned)ceil_div(nc, 8), 256, 0, st>>>(part_c, sc, nc, (int)d, dc, dc_bf16);
  TT_LAUNCH_OK("combine_partials_kernel");
  return TT_OK;
}

}  // namespace tt
