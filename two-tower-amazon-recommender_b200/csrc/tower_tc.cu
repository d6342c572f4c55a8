// K1+K2 fused (bf16 precision) -- a whole two-layer tower in one kernel, forward and backward.
//
//   forward : gather/pool 128 batch rows (fp32 tables, sequential fp32 sums) straight into the
//             swizzled shared-memory A tile, D1 = x W1 on tcgen05 (TMEM), epilogue +b1, relu ->
//             bf16 h back into shared memory (over the dead W1 tile) and to HBM for the backward,
//             D2 = h W2 on tcgen05, epilogue +b2 -> bf16 y.  W1/W2 arrive by TMA in the Keras
//             [in, out] layout and are read as MN-major B operands: no transposed copy exists.
//   backward: dy = ordered sum of the fp32 split partials of the loss backward (rounded to bf16
//             once, db2 taken before rounding), then four GEMMs on the same 128 rows:
//               dh_pre = dy W2^T           (A = dy K-major,  B = W2 K-major)
//               dW2_p  = h^T dy            (A = h MN-major,  B = dy MN-major)
//               dx     = dh W1^T           (A = dh K-major,  B = W1 K-major)
//               dW1_p  = x^T dh            (A = x MN-major,  B = dh MN-major)
//             with dh = dh_pre * (h > 0) produced by the epilogue warps directly into shared memory,
//             where ONE copy serves as the K-major A of the third and the MN-major B of the fourth.
//             Shared memory and TMEM are re-used as operands die (see the region table below).
//
// One CTA = 128 rows of one tower, 256 threads; blockIdx.y selects the tower so query and candidate
// towers share a launch.  Both kernels are latency-bound at the BASELINE sizes (a tower is
// 0.5 GFLOP forward): what matters is that the gather, four/two GEMMs and all epilogues cost ONE
// launch and no HBM round trips for x / h / dh.
#include "tc_common.cuh"
#include "gather_row.cuh"
#include "sparse_ws.cuh"

namespace tt {

constexpr int TW_BM = 128;
constexpr int TW_BLK = TW_BM * 128;          // bytes of one [128 rows][64 bf16] swizzled block
constexpr int TW_STAGE = TW_BM * 128;        // one [128 rows][32 fp32] swizzled staging tile of the fp32 TMA stores

struct TowerDev {
  FeatureParams feats;
  int64_t B;
  int d_in, d_hid, d_out;
  const float* b1;
  const float* b2;
  uint16_t* x;
  uint16_t* h;
  uint16_t* y;
  SparseWs prep;       // prep_on: the control warp also runs the sparse optimizer's prepare stage (hash insert of the
  int prep_on;         // ids) for this ID-only tower while the workers gather -- no separate launch, no stream join
};

struct TowerFwdArgs {
  CUtensorMap tmW1[TT_MAX_TOWERS];   // W1 [d_in, d_hid]: box {64, d_in}
  CUtensorMap tmW2[TT_MAX_TOWERS];   // W2 [d_hid, d_out]: box {64, d_hid}
  CUtensorMap tmH[TT_MAX_TOWERS];    // h [B, d_hid] (store): box {64, 128}
  CUtensorMap tmY[TT_MAX_TOWERS];    // y [B, d_out] (store): box {64, 128}
  CUtensorMap tmX[TT_MAX_TOWERS];    // x [B, d_in] (load; only when the tower input is given instead of gathered)
  TowerDev t[TT_MAX_TOWERS];
  int* fault;
  long long* trace;
};

struct TowerBwdDev {
  int64_t B;
  int d_in, d_hid, d_out;
  const float* dy_parts;
  int dy_splits;
  float* dx;
  float* dw1_parts;
  float* dw2_parts;
  float* db1_parts;
  float* db2_parts;
};

struct TowerBwdArgs {
  CUtensorMap tmX[TT_MAX_TOWERS];    // x [B, d_in]:  box {64, 128}
  CUtensorMap tmH[TT_MAX_TOWERS];    // h [B, d_hid]: box {64, 128}
  CUtensorMap tmW2[TT_MAX_TOWERS];   // W2 [d_hid, d_out] read K-major (N = d_hid rows): box {64, d_hid}
  CUtensorMap tmW1[TT_MAX_TOWERS];   // W1 [d_in, d_hid]  read K-major (N = d_in rows):  box {64, d_in}
  CUtensorMap tmDX[TT_MAX_TOWERS];   // stores, fp32, box {32 columns, 32 rows}: dx [B, d_in]
  CUtensorMap tmDW1[TT_MAX_TOWERS];  //   dW1 partials viewed as [P * d_in, d_hid]
  CUtensorMap tmDW2[TT_MAX_TOWERS];  //   dW2 partials viewed as [P * d_hid, d_out]
  TowerBwdDev t[TT_MAX_TOWERS];
  long long* trace;
};

// Phase stamps for tuning (tools/trace_tower.py): 16 globaltimer slots per CTA, fwd CTAs then bwd CTAs.
static long long* g_tower_trace = nullptr;
TT_TL_DEFINE(set_timeline_tower)
__device__ __forceinline__ long long tw_now() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TW_STAMP(i)                                                                              \
  do {                                                                                           \
    if (a.trace && threadIdx.x == 0) a.trace[16 * (blockIdx.y * gridDim.x + blockIdx.x) + (i)] = tw_now(); \
  } while (0)

__host__ __device__ inline int imax(int a, int b) { return a > b ? a : b; }

// 16-byte-chunk store of 8 bf16 into a K-major swizzled tile made of [128][64] blocks
__device__ __forceinline__ void st_tile_chunk(uint8_t* tile, int r, int col, uint4 v) {
  *reinterpret_cast<uint4*>(tile + (col >> 6) * TW_BLK + sw128_offset(r, (col & 63) >> 3)) = v;
}

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
struct FwdLayout {
  int x_bytes, r1_bytes, w1_bytes, w2_bytes, off_r1, off_w2, off_tail, total;
};
__host__ __device__ inline FwdLayout fwd_layout(int d_in, int d_hid, int d_out) {
  FwdLayout L;
  L.x_bytes = TW_BM * d_in * 2;
  L.w1_bytes = d_in * d_hid * 2;
  L.w2_bytes = d_hid * d_out * 2;
  L.r1_bytes = imax(L.w1_bytes, TW_BM * d_hid * 2);   // W1, then h
  L.off_r1 = L.x_bytes;
  L.off_w2 = L.off_r1 + L.r1_bytes;
  L.off_tail = L.off_w2 + L.w2_bytes;
  L.total = L.off_tail + 128 + (d_hid + d_out) * 4;
  return L;
}

// 8 worker warps (gather, epilogues) + 1 control warp (TMA, MMA issue) [+ 1 dedup warp]: the hand-offs are mbarriers, so the two
// halves of the hidden layer pipeline -- epilogue of half 0 under the MMAs of half 1, the second GEMM's first
// K-half under the epilogue of half 1 -- instead of block-wide barriers between whole phases.
constexpr int TWF_THREADS = 288;
constexpr int TWF_THREADS_FWD = 320;   // + the dedup warp of the forward

__global__ void __launch_bounds__(TWF_THREADS_FWD, 1)
tower_mlp2_fwd_kernel(const __grid_constant__ TowerFwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int tw = blockIdx.y;
  const TowerDev& P = a.t[tw];
  const int64_t m0 = (int64_t)blockIdx.x * TW_BM;
  if (m0 >= P.B) return;
  const int d_in = P.d_in, d_hid = P.d_hid, d_out = P.d_out;
  const FwdLayout L = fwd_layout(d_in, d_hid, d_out);
  TW_STAMP(0);
  long long* const tl = g_tl;
  uint8_t* sX = smem;
  uint8_t* sR1 = smem + L.off_r1;                 // W1 (MN-major chunks [d_in][64]) then h (K-major blocks)
  uint8_t* sW2 = smem + L.off_w2;                 // MN-major chunks [d_hid][64]
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + L.off_tail);
  uint64_t* x_ready = w_full + 1;
  uint64_t* mma1_done = w_full + 2;               // [2]
  uint64_t* h_ready = w_full + 4;                 // [2]
  uint64_t* mma2_done = w_full + 6;
  uint64_t* y_ready = w_full + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 8);
  float* sB1 = reinterpret_cast<float*>(smem + L.off_tail + 128);
  float* sB2 = sB1 + d_hid;
  const int NH = d_hid / 128;                     // halves of the hidden layer (128 columns each)
  const bool x_by_tma = P.feats.n == 0;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.tmW1[tw]);
    tma_prefetch_desc(&a.tmW2[tw]);
    tma_prefetch_desc(&a.tmH[tw]);
    tma_prefetch_desc(&a.tmY[tw]);
    mbar_init(w_full, 1);
    mbar_init(x_ready, 8 + (x_by_tma ? 1 : 0));   // one arrival per worker warp (+ the TMA transaction)
    mbar_init(mma1_done, 1); mbar_init(mma1_done + 1, 1);
    mbar_init(h_ready, 8); mbar_init(h_ready + 1, 8);
    mbar_init(mma2_done, 1);
    mbar_init(y_ready, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                       // the prologue above overlapped the tail of the previous kernel in the stream
  pdl_launch_dependents();
  tl_mark(tl, 0, true);
  TW_STAMP(1);

  if (warp == 8) {
    // ================= control warp: one elected lane issues, the warp stays converged around the waits
    if (elect_one_sync()) {                         // weights: L2-resident after the first CTA
      mbar_arrive_expect_tx(w_full, L.w1_bytes + L.w2_bytes);
      for (int j = 0; j < d_hid / 64; ++j) tma_load_2d(sR1 + j * d_in * 128, &a.tmW1[tw], w_full, 64 * j, 0);
      for (int j = 0; j < d_out / 64; ++j) tma_load_2d(sW2 + j * d_hid * 128, &a.tmW2[tw], w_full, 64 * j, 0);
      if (x_by_tma) {                               // tower input given (row-sharded lookup already exchanged): TMA it in
        mbar_arrive_expect_tx(x_ready, L.x_bytes);
        for (int j = 0; j < d_in / 64; ++j) tma_load_2d(sX + j * TW_BLK, &a.tmX[tw], x_ready, 64 * j, (int)m0);
      }
    }
    __syncwarp();
    mbar_wait(x_ready, 0);
    mbar_wait(w_full, 0);
    tc_fence_after();
    if (elect_one_sync()) {
      // D1[:, 128 hh .. 128 hh + 128) = x W1[:, same columns]: one commit per half
      const int n1 = d_hid / NH;
      const uint32_t idesc = umma_idesc_bf16(TW_BM, n1, 0, 1);
      for (int hh = 0; hh < NH; ++hh) {
        const uint64_t db0 = umma_desc_mn_sw128(smem_u32(sR1 + 2 * hh * d_in * 128), d_in * 128);
        for (int kk = 0; kk < d_in / 16; ++kk) {
          const uint64_t da = umma_desc_k_sw128(smem_u32(sX + (kk >> 2) * TW_BLK)) + 2 * (kk & 3);
          umma_bf16_ss(tmem_base + 128 * hh, da, db0 + 128 * kk, idesc, kk != 0);
        }
        umma_commit(mma1_done + hh);
      }
    }
    __syncwarp();
    for (int hh = 0; hh < NH; ++hh) {
      mbar_wait(h_ready + hh, 0);                   // the bf16 h columns of this half are in shared memory
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t idesc = umma_idesc_bf16(TW_BM, d_out, 0, 1);
        const uint64_t db0 = umma_desc_mn_sw128(smem_u32(sW2), d_hid * 128);
        for (int kk = 8 * hh; kk < 8 * hh + 8; ++kk) {
          const uint64_t da = umma_desc_k_sw128(smem_u32(sR1 + (kk >> 2) * TW_BLK)) + 2 * (kk & 3);
          umma_bf16_ss(tmem_base + 256, da, db0 + 128 * kk, idesc, kk != 0);
        }
        if (hh == NH - 1) umma_commit(mma2_done);
        for (int j = 2 * hh; j < 2 * hh + 2; ++j) tma_store_2d(&a.tmH[tw], sR1 + j * TW_BLK, 64 * j, (int)m0);
        tma_store_commit();
      }
      __syncwarp();
    }
    mbar_wait(y_ready, 0);
    if (elect_one_sync()) {
      for (int j = 0; j < d_out / 64; ++j) tma_store_2d(&a.tmY[tw], sX + j * TW_BLK, 64 * j, (int)m0);
      tma_store_commit();
      tma_store_wait_read<0>();                     // shared memory must outlive the reads of every bulk store
    }
    __syncwarp();
  } else if (warp == 9) {
    // ================= dedup warp: the sparse optimizer's prepare stage for this CTA's 128 ids (what
    // tt_optimizer_prepare_sparse does for a whole id list), hidden under the gather and the GEMMs; every lane keeps
    // the probe round trips of its 4 entries in flight together (sparse_ws.cuh).
    if (P.prep_on) {
      constexpr int U = TW_BM / 32;
      const tt_feature& ft = P.feats.f[0];
      const SparseWs& ws = P.prep;
      int64_t id[U], j[U];
      bool pend[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        j[u] = m0 + lane + 32 * u;
        id[u] = j[u] < P.B ? __ldg(ft.values + j[u]) : -2;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        pend[u] = id[u] >= 0 && id[u] < ft.vocab;
        if (!pend[u] && id[u] != -2) ws.hpos[j[u]] = -1;               // out of range: dropped by the optimizer as well
      }
      if (a.trace && lane == 0) a.trace[16 * (blockIdx.y * gridDim.x + blockIdx.x) + 9] = tw_now();
      sparse_prepare_entries<U>(ws, j, id, pend);
      if (a.trace && lane == 0) a.trace[16 * (blockIdx.y * gridDim.x + blockIdx.x) + 10] = tw_now();
    }
  } else {
    // ================= worker warps
    for (int i = threadIdx.x; i < d_hid; i += 256) sB1[i] = P.b1[i];
    for (int i = threadIdx.x; i < d_out; i += 256) sB2[i] = P.b2[i];

    // ---- gather / pool the 128 rows: lane l owns columns 4l..4l+3 of the row (d_in == 128)
    if (!x_by_tma) {
      const int nchunks = d_in >> 2;
      const int col = 4 * lane;
      auto put_row = [&](int r, int64_t b, const float4& v) {
        if (lane < nchunks) {
          const uint2 pk = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
          *reinterpret_cast<uint2*>(sX + (col >> 6) * TW_BLK + sw128_offset(r, (col & 63) >> 3) + (lane & 1) * 8) = pk;
          if (b < P.B && P.x) *reinterpret_cast<uint2*>(P.x + b * d_in + col) = pk;
        }
      };
      if (P.feats.n == 1 && P.feats.f[0].offsets == nullptr) {
        // ID-only tower: the warp's 16 ids in one load, then 16 independent row loads in flight per lane
        const tt_feature& ft = P.feats.f[0];
        int64_t my_id = -1;
        if (lane < TW_BM / 8) {
          const int64_t b = m0 + warp + 8 * lane;
          if (b < P.B) {
            my_id = __ldg(ft.values + b);
            if (my_id < 0 || my_id >= ft.vocab) {
              if (a.fault && my_id != -1) atomicExch(a.fault, 1);
              my_id = -1;
            }
          }
        }
        float4 v[TW_BM / 8];
#pragma unroll
        for (int i = 0; i < TW_BM / 8; ++i) {
          const int64_t id = __shfl_sync(0xffffffffu, my_id, i);
          v[i] = (id >= 0 && lane < nchunks) ? ldg_feature_row(ft, id, d_in, lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < TW_BM / 8; ++i) put_row(warp + 8 * i, m0 + warp + 8 * i, v[i]);
      } else {
#pragma unroll 4
        for (int i = 0; i < TW_BM / 8; ++i) {
          const int r = warp + 8 * i;
          const int64_t b = m0 + r;
          float4 v[1];
          v[0] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (b < P.B) gather_row<1>(P.feats, b, d_in, lane, nchunks, a.fault, v);
          put_row(r, b, v[0]);
        }
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(x_ready);            // also publishes the bias vectors to the other workers (see below)
    TW_STAMP(2);

    const int q = warp & 3, hf = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);

    // ---- epilogue 1, per half: h = relu(D1 + b1) -> bf16 into shared memory (over the W1 chunks that half has
    // consumed); the swizzled tile is both the A operand of the second GEMM and the source of the TMA store of h.
    // The bias vectors written by the other warps are visible here: every worker arrived on x_ready (release), the
    // control thread acquired it before issuing the MMAs whose commit this wait acquires.
    for (int hh = 0; hh < NH; ++hh) {
      mbar_wait(mma1_done + hh, 0);
      if (hh == 0) TW_STAMP(3);
      tc_fence_after();
      const int c_base = 128 * hh + 64 * hf;
      uint32_t rr[64];
      tmem_ld32(lane_addr + c_base, rr);
      tmem_ld32(lane_addr + c_base + 32, rr + 32);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 ba = *reinterpret_cast<const float4*>(sB1 + c_base + g * 8);
        const float4 bb = *reinterpret_cast<const float4*>(sB1 + c_base + g * 8 + 4);
        const float bias[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(__uint_as_float(rr[g * 8 + j]) + bias[j], 0.f);
        st_tile_chunk(sR1, r, c_base + g * 8,
                      make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7])));
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(h_ready + hh);
      TW_STAMP(4 + hh);
    }

    // ---- epilogue 2: y = D2 + b2 -> bf16, staged in the (dead) x tile and written by one TMA store
    mbar_wait(mma2_done, 0);
    TW_STAMP(6);
    tc_fence_after();
    {
      const int ncol = d_out / 2;                   // 64 or 32 columns per warp
      const int c_base = hf * ncol;
      uint32_t rr[64];
      tmem_ld32(lane_addr + 256 + c_base, rr);
      if (ncol > 32) tmem_ld32(lane_addr + 256 + c_base + 32, rr + 32);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        if (g * 8 < ncol) {
          const float4 ba = *reinterpret_cast<const float4*>(sB2 + c_base + g * 8);
          const float4 bb = *reinterpret_cast<const float4*>(sB2 + c_base + g * 8 + 4);
          const float bias[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(rr[g * 8 + j]) + bias[j];
          st_tile_chunk(sX, r, c_base + g * 8,
                        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7])));
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(y_ready);
    TW_STAMP(7);
  }
  __syncthreads();                                  // every tcgen05.ld and every bulk-store read has finished
  TW_STAMP(8);
  tl_mark(tl, 0, false);
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------
// Shared-memory regions and their successive tenants (cfg2: 32 + 64 + 64 + 32 KB):
//   R_dy : dy tile [128][d_out] (K-major blocks)      -> second half of the W1 k-blocks
//   R_h  : h tile  [128][d_hid]                        -> x tile [128][d_in]
//   R_w2 : W2 k-blocks [d_hid][64] x d_out/64          -> dh tile [128][d_hid]
//   R_w1 : first half of the W1 k-blocks [d_in][64]
// TMEM columns: dh_pre [0, d_hid) -> dx [0, d_in); dW2 [256, 256 + d_hid/128 * d_out);
//               dW1 [d_in, d_in + d_hid) once dW2 has been drained.
struct BwdTLayout {
  int dy_bytes, h_bytes, w2r_bytes, w1blk, w1lo_blocks, nk1;
  int off_h, off_w2, off_w1, off_tail, total;
};
__host__ __device__ inline BwdTLayout bwdt_layout(int d_in, int d_hid, int d_out) {
  BwdTLayout L;
  L.dy_bytes = TW_BM * d_out * 2;
  L.h_bytes = TW_BM * d_hid * 2;
  L.w2r_bytes = imax(d_hid * d_out * 2, TW_BM * d_hid * 2);
  L.w1blk = d_in * 128;
  L.nk1 = d_hid / 64;
  L.w1lo_blocks = L.nk1 - L.dy_bytes / L.w1blk < 0 ? 0 : L.nk1 - L.dy_bytes / L.w1blk;   // blocks that do not fit R_dy
  if (L.w1lo_blocks < L.nk1 / 2) L.w1lo_blocks = L.nk1 / 2;
  L.off_h = L.dy_bytes;
  L.off_w2 = L.off_h + L.h_bytes;
  L.off_w1 = L.off_w2 + L.w2r_bytes;
  L.off_tail = L.off_w1 + L.w1lo_blocks * L.w1blk;
  // tail: barriers (128 B) + eight 4 KB per-warp transposing tiles of the fp32 outputs (the db2 scratch lives there first)
  L.total = L.off_tail + 1024 + 2 * TW_STAGE;
  return L;
}

// Same warp roles as the forward: 8 worker warps (dy tile, epilogues, drains) + 1 control warp (TMA, MMA issue).
// Order of the tensor-pipe work and what each piece waits for:
//   a  dh_pre = dy W2^T        dy tile built, W2 landed                      (h still in flight)
//   b  dW2    = h^T dy         h landed
//   c  dx     = dh W1^T        per 128-column half of dh as the epilogue stages it
//   d  dW1    = x^T dh         half 0 into the dead dh_pre columns at once, half 1 once the first 128 dW2
//                              columns have been drained (TMEM holds 512 of the 640 columns the outputs need)
// The workers drain dW2, dx, dW1 through TWO 4 KB staging tiles per warp (the second in the dead half of the
// h region), so a TMA store reads one tile while the warp fills the other.
__global__ void __launch_bounds__(TWF_THREADS, 1)
tower_mlp2_bwd_kernel(const __grid_constant__ TowerBwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int tw = blockIdx.y;
  const TowerBwdDev& P = a.t[tw];
  const int64_t m0 = (int64_t)blockIdx.x * TW_BM;
  if (m0 >= P.B) return;
  const int d_in = P.d_in, d_hid = P.d_hid, d_out = P.d_out;
  const BwdTLayout L = bwdt_layout(d_in, d_hid, d_out);
  TW_STAMP(0);
  long long* const tl = g_tl;
  uint8_t* rDY = smem;
  uint8_t* rH = smem + L.off_h;
  uint8_t* rW2 = smem + L.off_w2;
  uint8_t* rW1 = smem + L.off_w1;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_tail);
  uint64_t* w2_full = bars + 0;
  uint64_t* h_full = bars + 1;
  uint64_t* w1lo_full = bars + 2;
  uint64_t* w1hi_full = bars + 3;
  uint64_t* x_full = bars + 4;
  uint64_t* dy_ready = bars + 5;
  uint64_t* mma_a_done = bars + 6;
  uint64_t* mma_b_done = bars + 7;
  uint64_t* mma_c_done = bars + 8;
  uint64_t* mma_d_done = bars + 9;                 // [2]
  uint64_t* dh_ready = bars + 11;                  // [2]
  uint64_t* dw2_lo_drained = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  uint8_t* stage = smem + L.off_tail + 1024;       // [8 warps][32 rows][32 fp32]
  float* s_db2 = reinterpret_cast<float*>(stage);  // [8][d_out], dead before the first store is staged
  const int p = blockIdx.x;                        // partial index of this 128-row slice
  const int n_hi = L.nk1 - L.w1lo_blocks;          // W1 k-blocks that move into R_dy later
  const int NH = d_hid / 128;                      // 128-column halves of dh
  const int C2 = (d_hid / 128) * d_out;            // concatenated columns of the dW2 tiles
  const int x_bytes = TW_BM * d_in * 2;
  const bool two_stages = L.h_bytes >= x_bytes + 8 * 4096;   // room for a second staging tile per warp behind x

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.tmX[tw]); tma_prefetch_desc(&a.tmH[tw]);
    tma_prefetch_desc(&a.tmW2[tw]); tma_prefetch_desc(&a.tmW1[tw]);
    tma_prefetch_desc(&a.tmDX[tw]); tma_prefetch_desc(&a.tmDW1[tw]); tma_prefetch_desc(&a.tmDW2[tw]);
    for (int i = 0; i < 5; ++i) mbar_init(bars + i, 1);
    mbar_init(dy_ready, 8);
    for (int i = 6; i < 11; ++i) mbar_init(bars + i, 1);
    mbar_init(dh_ready, 8); mbar_init(dh_ready + 1, 8);
    mbar_init(dw2_lo_drained, (C2 / 2 >= 128) ? 4 : 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                       // the prologue above overlapped the tail of the previous kernel in the stream
  pdl_launch_dependents();
  tl_mark(tl, 4, true);
  TW_STAMP(1);

  if (warp == 8) {
    // ================= control warp
    if (elect_one_sync()) {                         // W2 first: the first GEMM needs only W2 and the dy tile
      mbar_arrive_expect_tx(w2_full, d_hid * d_out * 2);
      for (int j = 0; j < d_out / 64; ++j) tma_load_2d(rW2 + j * d_hid * 128, &a.tmW2[tw], w2_full, 64 * j, 0);
      mbar_arrive_expect_tx(h_full, L.h_bytes);
      for (int j = 0; j < d_hid / 64; ++j) tma_load_2d(rH + j * TW_BLK, &a.tmH[tw], h_full, 64 * j, (int)m0);
      mbar_arrive_expect_tx(w1lo_full, L.w1lo_blocks * L.w1blk);
      for (int j = 0; j < L.w1lo_blocks; ++j) tma_load_2d(rW1 + j * L.w1blk, &a.tmW1[tw], w1lo_full, 64 * j, 0);
    }
    __syncwarp();
    mbar_wait(dy_ready, 0);
    mbar_wait(w2_full, 0);
    tc_fence_after();
    if (elect_one_sync()) {  // a: dh_pre[128, d_hid] = dy W2^T
      const uint32_t idesc = umma_idesc_bf16(TW_BM, d_hid, 0, 0);
      for (int kk = 0; kk < d_out / 16; ++kk) {
        const uint64_t da = umma_desc_k_sw128(smem_u32(rDY + (kk >> 2) * TW_BLK)) + 2 * (kk & 3);
        const uint64_t db = umma_desc_k_sw128(smem_u32(rW2 + (kk >> 2) * d_hid * 128)) + 2 * (kk & 3);
        umma_bf16_ss(tmem_base, da, db, idesc, kk != 0);
      }
      umma_commit(mma_a_done);
    }
    __syncwarp();
    mbar_wait(h_full, 0);
    tc_fence_after();
    if (elect_one_sync()) {  // b: dW2[m-th 128 rows of d_hid, d_out] = h^T dy  (K = the 128 batch rows)
      const uint32_t idesc = umma_idesc_bf16(128, d_out, 1, 1);
      const uint64_t db0 = umma_desc_mn_sw128(smem_u32(rDY), TW_BLK);
      for (int m = 0; m < d_hid / 128; ++m) {
        const uint64_t da0 = umma_desc_mn_sw128(smem_u32(rH + 2 * m * TW_BLK), TW_BLK);
        for (int kk = 0; kk < TW_BM / 16; ++kk)
          umma_bf16_ss(tmem_base + 256 + m * d_out, da0 + 128 * kk, db0 + 128 * kk, idesc, kk != 0);
      }
      umma_commit(mma_b_done);
    }
    __syncwarp();
    mbar_wait(mma_b_done, 0);
    if (n_hi > 0 && elect_one_sync()) {             // dy is dead once both GEMMs are done: bring in the rest of W1
      mbar_arrive_expect_tx(w1hi_full, n_hi * L.w1blk);
      for (int j = 0; j < n_hi; ++j) tma_load_2d(rDY + j * L.w1blk, &a.tmW1[tw], w1hi_full, 64 * (L.w1lo_blocks + j), 0);
    }
    __syncwarp();
    mbar_wait(w1lo_full, 0);
    bool hi_seen = n_hi == 0;
    for (int hh = 0; hh < NH; ++hh) {
      mbar_wait(dh_ready + hh, 0);                  // dh columns [128 hh, 128 hh + 128) staged; that half of h is dead
      if (!hi_seen && 2 * (hh + 1) > L.w1lo_blocks) { mbar_wait(w1hi_full, 0); hi_seen = true; }
      tc_fence_after();
      if (elect_one_sync()) {
        if (hh == 0) {                              // x takes the place of the first h blocks (d_in <= 128 columns)
          mbar_arrive_expect_tx(x_full, x_bytes);
          for (int j = 0; j < d_in / 64; ++j) tma_load_2d(rH + j * TW_BLK, &a.tmX[tw], x_full, 64 * j, (int)m0);
        }
        // c: dx[128, d_in] = dh W1^T, the K steps of this half
        const uint32_t idesc = umma_idesc_bf16(TW_BM, d_in, 0, 0);
        for (int kk = 8 * hh; kk < 8 * hh + 8; ++kk) {
          const int kb = kk >> 2;
          const uint8_t* wblk = kb < L.w1lo_blocks ? rW1 + kb * L.w1blk : rDY + (kb - L.w1lo_blocks) * L.w1blk;
          const uint64_t da = umma_desc_k_sw128(smem_u32(rW2 + kb * TW_BLK)) + 2 * (kk & 3);
          const uint64_t db = umma_desc_k_sw128(smem_u32(wblk)) + 2 * (kk & 3);
          umma_bf16_ss(tmem_base, da, db, idesc, kk != 0);
        }
        if (hh == NH - 1) umma_commit(mma_c_done);
      }
      __syncwarp();
    }
    mbar_wait(x_full, 0);
    for (int hh = 0; hh < NH; ++hh) {
      if (hh == 1) mbar_wait(dw2_lo_drained, 0);    // the columns dW1's second half lands on
      tc_fence_after();
      if (elect_one_sync()) {  // d: dW1[d_in, 128 hh .. +128) = x^T dh  (K = the 128 batch rows)
        const uint32_t idesc = umma_idesc_bf16(128, d_hid / NH, 1, 1);
        const uint64_t da0 = umma_desc_mn_sw128(smem_u32(rH), TW_BLK);
        const uint64_t db0 = umma_desc_mn_sw128(smem_u32(rW2 + 2 * hh * TW_BLK), TW_BLK);
        for (int kk = 0; kk < TW_BM / 16; ++kk)
          umma_bf16_ss(tmem_base + d_in + 128 * hh, da0 + 128 * kk, db0 + 128 * kk, idesc, kk != 0);
        umma_commit(mma_d_done + hh);
      }
      __syncwarp();
    }
  } else {
    // ================= worker warps
    // ---- dy = ordered sum of the split partials; db2 from the fp32 sums; bf16 tile into R_dy.
    // 8 rows x 2 splits = 16 independent 128-bit loads in flight per lane (the phase is L2-bandwidth-bound)
    {
      const int nchunks = d_out >> 2;                // float4 chunks per row (<= 32)
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const int col = 4 * lane;
      const bool lane_ok = lane < nchunks;
      const int S = P.dy_splits;
#pragma unroll 1
      for (int i0 = 0; i0 < TW_BM / 8; i0 += 8) {
        float4 v[8], u[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t b = m0 + warp + 8 * (i0 + i);
          const bool ok = lane_ok && b < P.B;
          v[i] = ok ? __ldg(reinterpret_cast<const float4*>(P.dy_parts + b * d_out) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
          u[i] = (ok && S > 1) ? __ldg(reinterpret_cast<const float4*>(P.dy_parts + (P.B + b) * d_out) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = warp + 8 * (i0 + i);
          const int64_t b = m0 + r;
          v[i].x += u[i].x; v[i].y += u[i].y; v[i].z += u[i].z; v[i].w += u[i].w;
          if (lane_ok && b < P.B)
            for (int sp = 2; sp < S; ++sp) {
              const float4 w = __ldg(reinterpret_cast<const float4*>(P.dy_parts + ((int64_t)sp * P.B + b) * d_out) + lane);
              v[i].x += w.x; v[i].y += w.y; v[i].z += w.z; v[i].w += w.w;
            }
          acc.x += v[i].x; acc.y += v[i].y; acc.z += v[i].z; acc.w += v[i].w;
          if (lane_ok)
            *reinterpret_cast<uint2*>(rDY + (col >> 6) * TW_BLK + sw128_offset(r, (col & 63) >> 3) + (lane & 1) * 8) =
                make_uint2(pack_bf16x2(v[i].x, v[i].y), pack_bf16x2(v[i].z, v[i].w));
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(dy_ready);
      if (lane_ok) *reinterpret_cast<float4*>(s_db2 + warp * d_out + col) = acc;
    }
    TW_STAMP(2);                                     // dy tile built
    asm volatile("bar.sync 1, 256;" ::: "memory");   // the workers only: s_db2 complete
    if (threadIdx.x < d_out) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += s_db2[w * d_out + threadIdx.x];
      P.db2_parts[(int64_t)p * d_out + threadIdx.x] = s;
    }

    const int q = warp & 3, hf = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);

    // ---- epilogue a, per half: dh = dh_pre * (h > 0) -> bf16 tile over the dead W2
    mbar_wait(mma_a_done, 0);
    TW_STAMP(3);
    mbar_wait(h_full, 0);
    tc_fence_after();
    for (int hh = 0; hh < NH; ++hh) {
      const int c_base = 128 * hh + 64 * hf;
      uint32_t rr[64];
      tmem_ld32(lane_addr + c_base, rr);
      tmem_ld32(lane_addr + c_base + 32, rr + 32);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int col = c_base + g * 8;
        const uint4 hv = *reinterpret_cast<const uint4*>(rH + (col >> 6) * TW_BLK + sw128_offset(r, (col & 63) >> 3));
        const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
        float v[8];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float lo = __uint_as_float(hw[t] << 16), hi = __uint_as_float(hw[t] & 0xffff0000u);
          v[2 * t] = lo > 0.f ? __uint_as_float(rr[g * 8 + 2 * t]) : 0.f;
          v[2 * t + 1] = hi > 0.f ? __uint_as_float(rr[g * 8 + 2 * t + 1]) : 0.f;
        }
        st_tile_chunk(rW2, r, col, make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7])));
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dh_ready + hh);
    }
    TW_STAMP(4);

    // fp32 outputs leave through shared memory: every warp stages its [32 rows x 32 columns] piece (row =
    // TMEM lane, 128B-swizzled, 4 KB) and its first lane issues a TMA store of it: full-line writes by the
    // copy engine instead of 32 scattered 16-byte stores per warp instruction, and no cross-warp barrier.
    // Two tiles per warp alternate when the h region has room behind x.  (Measured: the drains run at the chip's
    // L2 write rate, ~6 TB/s for the 41 MB of dx + dW partials of a cfg2 step; transposing through the tile and
    // writing with coalesced 128-bit stores instead of TMA was 25 % slower.)
    uint8_t* const stage_a = stage + warp * 4096;
    uint8_t* const stage_b = two_stages ? rH + x_bytes + warp * 4096 : stage_a;
    int flip = 0;
    auto store_tile = [&](uint32_t taddr, const CUtensorMap* tm, int c_inner, int c_outer) {
      uint32_t rr[32];
      tmem_ld32(taddr, rr);
      if (lane == 0) {                               // the tile about to be overwritten has been read by its store
        if (two_stages) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
      }
      __syncwarp();
      tmem_ld_wait();
      uint8_t* const my_stage = flip ? stage_b : stage_a;
      flip ^= 1;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        *reinterpret_cast<uint4*>(my_stage + sw128_offset(lane, g)) = make_uint4(rr[4 * g], rr[4 * g + 1], rr[4 * g + 2], rr[4 * g + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { tma_store_2d(tm, my_stage, c_inner, c_outer + 32 * q); tma_store_commit(); }
    };

    // every worker is past its reads of s_db2 and of h, and dW2 is complete: the staging tiles are free
    mbar_wait(dh_ready + NH - 1, 0);
    mbar_wait(mma_b_done, 0);
    tc_fence_after();
    // ---- dW2 partial -> HBM (rows = d_hid index, TMEM lanes), while dx runs on the tensor pipe
#pragma unroll 1
    for (int cc = hf * (C2 / 2); cc < (hf + 1) * (C2 / 2); cc += 32) {
      const int m = cc / d_out, n0 = cc % d_out;
      store_tile(lane_addr + 256 + cc, &a.tmDW2[tw], n0, p * d_hid + m * 128);
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0 && hf * (C2 / 2) < 128) mbar_arrive(dw2_lo_drained);
    TW_STAMP(5);
    // ---- db1 partial = column sums of the bf16 dh tile: a thread owns a column pair and half of the rows (the
    // two halves differ in row bit 2, so their swizzled words fall on different banks), two independent chains
    if (warp * 32 < d_hid) {
      const int col = 2 * (warp * 16 + (lane & 15)), rh = lane >> 4;
      const uint8_t* base = rW2 + (col >> 6) * TW_BLK + (col & 7) * 2;
      const int ch = (col & 63) >> 3;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
      for (int i = 0; i < TW_BM / 2; ++i) {
        const int row = (i & 3) + 8 * (i >> 2) + 4 * rh;
        const uint32_t w = *reinterpret_cast<const uint32_t*>(base + sw128_offset(row, ch));
        s0 += __uint_as_float(w << 16);
        s1 += __uint_as_float(w & 0xffff0000u);
      }
      s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
      if (rh == 0) *reinterpret_cast<float2*>(P.db1_parts + (int64_t)p * d_hid + col) = make_float2(s0, s1);
    }

    // ---- dx -> HBM fp32 (the embedding-row gradient); rows past the batch are clipped by the tensor map
    mbar_wait(mma_c_done, 0);
    TW_STAMP(6);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = hf * (d_in / 2); c0 < (hf + 1) * (d_in / 2); c0 += 32)
      store_tile(lane_addr + c0, &a.tmDX[tw], c0, (int)m0);
    TW_STAMP(7);
    // ---- dW1 partial -> HBM (rows = d_in index)
    mbar_wait(mma_d_done + (NH == 2 ? hf : 0), 0);
    TW_STAMP(8);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = hf * (d_hid / 2); c0 < (hf + 1) * (d_hid / 2); c0 += 32)
      store_tile(lane_addr + d_in + c0, &a.tmDW1[tw], c0, p * d_in);
    TW_STAMP(9);
    if (lane == 0) tma_store_wait_read<0>();
    __syncwarp();
    tc_fence_before();
  }
  __syncthreads();
  TW_STAMP(10);
  tl_mark(tl, 4, false);
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- host ----------------------------------------------------------------------------
static bool mlp2_supported(int d_in, int d_hid, int d_out) {
  if (d_in != 128) return false;                           // M of the dW1 GEMM
  if (d_hid != 128 && d_hid != 256) return false;          // M tiles of the dW2 GEMM
  if (d_out != 64 && d_out != 128) return false;
  const BwdTLayout L = bwdt_layout(d_in, d_hid, d_out);
  if ((L.nk1 - L.w1lo_blocks) * L.w1blk > L.dy_bytes) return false;
  return fwd_layout(d_in, d_hid, d_out).total <= 227 * 1024 && L.total <= 227 * 1024;
}

static int check_tower(const char* fn, const tt_tower_mlp2& s, bool bwd) {
  TT_REQUIRE(mlp2_supported(s.d_in, s.d_hid, s.d_out), "%s: unsupported tower shape %d-%d-%d (see tt_tower_mlp2_supported)", fn, s.d_in, s.d_hid, s.d_out);
  TT_REQUIRE(s.batch > 0 && s.batch < (1ll << 31), "%s: bad batch %lld", fn, (long long)s.batch);
  TT_REQUIRE(s.w1 && s.w2 && aligned16(s.w1) && aligned16(s.w2), "%s: weights null or unaligned", fn);
  TT_REQUIRE(s.x && s.h && aligned16(s.x) && aligned16(s.h), "%s: x / h buffers null or unaligned", fn);
  if (!bwd) {
    TT_REQUIRE(s.b1 && s.b2, "%s: bias null", fn);
    TT_REQUIRE(s.y && aligned16(s.y), "%s: y null or unaligned", fn);
    TT_REQUIRE(s.num_feats >= 0 && s.num_feats <= TT_MAX_FEATURES, "%s: num_feats must be in [0, %d] (0: x is an input)", fn, TT_MAX_FEATURES);
    for (int i = 0; i < s.num_feats; ++i) {
      TT_REQUIRE(s.feats[i].table && (s.feats[i].shard_world >= 2 || aligned16(s.feats[i].table)) && s.feats[i].values && s.feats[i].vocab > 0, "%s: feature %d is incomplete", fn, i);
      TT_REQUIRE(s.feats[i].mode == TT_POOL_SUM || s.feats[i].mode == TT_POOL_MEAN, "%s: feature %d bad pooling mode", fn, i);
    }
  } else {
    TT_REQUIRE(s.dy_parts && aligned16(s.dy_parts) && s.dy_splits >= 1, "%s: dy_parts null/unaligned or dy_splits < 1", fn);
    TT_REQUIRE(s.dx && s.dw1_parts && s.dw2_parts && s.db1_parts && s.db2_parts, "%s: null output buffer", fn);
    TT_REQUIRE(aligned16(s.dx) && aligned16(s.dw1_parts) && aligned16(s.dw2_parts), "%s: outputs must be 16-byte aligned", fn);
  }
  return TT_OK;
}

}  // namespace tt

using namespace tt;

// Debug hook: device buffer of 2 * 16 * 256 int64 receiving per-CTA phase stamps (globaltimer ns) of the fused
// tower forward (first half) and backward (second half); NULL = off.
extern "C" int tt_debug_tower_trace(long long* device_buf) {
  g_tower_trace = device_buf;
  return TT_OK;
}

extern "C" int32_t tt_tower_mlp2_supported(int32_t d_in, int32_t d_hid, int32_t d_out) {
  return mlp2_supported(d_in, d_hid, d_out) ? 1 : 0;
}

extern "C" int tt_tower_mlp2_fwd(const tt_tower_mlp2* towers, int32_t n, int32_t* fault, void* stream) {
  TT_REQUIRE(towers && n >= 1 && n <= TT_MAX_TOWERS, "tt_tower_mlp2_fwd: num_towers must be in [1, %d]", TT_MAX_TOWERS);
  static thread_local TowerFwdArgs args;     // 64-byte aligned storage for the tensor maps (copied at launch)
  int64_t max_b = 0;
  int smem = 0;
  for (int i = 0; i < n; ++i) {
    int rc = check_tower("tt_tower_mlp2_fwd", towers[i], false);
    if (rc) return rc;
    const tt_tower_mlp2& s = towers[i];
    TowerDev& d = args.t[i];
    d.feats.n = s.num_feats;
    for (int f = 0; f < s.num_feats; ++f) d.feats.f[f] = s.feats[f];
    d.B = s.batch; d.d_in = s.d_in; d.d_hid = s.d_hid; d.d_out = s.d_out;
    d.b1 = s.b1; d.b2 = s.b2; d.x = s.x; d.h = s.h; d.y = s.y;
    d.prep_on = 0;
    if (s.prepare_workspace) {
      TT_REQUIRE(s.num_feats == 1 && s.feats[0].offsets == nullptr && s.feats[0].shard_world < 2,
                 "tt_tower_mlp2_fwd: prepare_workspace needs a tower of exactly one unsharded ID feature");
      TT_REQUIRE(aligned16(s.prepare_workspace), "tt_tower_mlp2_fwd: prepare_workspace must be 16-byte aligned");
      // d of the workspace layout = the table's row width = d_in
      if (!ws_carve(s.batch, s.d_in, s.prepare_workspace, s.prepare_workspace_bytes, &d.prep))
        return set_error(TT_ERR_WORKSPACE, "tt_tower_mlp2_fwd: prepare_workspace of tower %d too small", i);
      d.prep_on = 1;
    }
    rc = make_tmap_bf16_2d(&args.tmW1[i], s.w1, (uint64_t)s.d_hid, (uint64_t)s.d_in, (uint64_t)s.d_hid * 2, 64, (uint32_t)s.d_in);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&args.tmW2[i], s.w2, (uint64_t)s.d_out, (uint64_t)s.d_hid, (uint64_t)s.d_out * 2, 64, (uint32_t)s.d_hid);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&args.tmH[i], s.h, (uint64_t)s.d_hid, (uint64_t)s.batch, (uint64_t)s.d_hid * 2, 64, TW_BM);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&args.tmY[i], s.y, (uint64_t)s.d_out, (uint64_t)s.batch, (uint64_t)s.d_out * 2, 64, TW_BM);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&args.tmX[i], s.x, (uint64_t)s.d_in, (uint64_t)s.batch, (uint64_t)s.d_in * 2, 64, TW_BM);
    if (rc) return rc;
    max_b = std::max<int64_t>(max_b, s.batch);
    smem = std::max(smem, fwd_layout(s.d_in, s.d_hid, s.d_out).total);
  }
  args.fault = fault;
  args.trace = g_tower_trace;
  cudaStream_t st = (cudaStream_t)stream;
  TT_CUDA_OK(cudaFuncSetAttribute(tower_mlp2_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid((unsigned)ceil_div(max_b, TW_BM), (unsigned)n);
  TT_PROF("tower_mlp2_fwd_kernel", st);
  TT_CUDA_OK(launch_pdl(tower_mlp2_fwd_kernel, grid, dim3(TWF_THREADS_FWD), (size_t)smem, st, args));
  TT_LAUNCH_OK("tower_mlp2_fwd_kernel");
  return TT_OK;
}

extern "C" int tt_tower_mlp2_bwd(const tt_tower_mlp2* towers, int32_t n, void* stream) {
  TT_REQUIRE(towers && n >= 1 && n <= TT_MAX_TOWERS, "tt_tower_mlp2_bwd: num_towers must be in [1, %d]", TT_MAX_TOWERS);
  static thread_local TowerBwdArgs args;
  int64_t max_b = 0;
  int smem = 0;
  for (int i = 0; i < n; ++i) {
    int rc = check_tower("tt_tower_mlp2_bwd", towers[i], true);
    if (rc) return rc;
    const tt_tower_mlp2& s = towers[i];
    TowerBwdDev& d = args.t[i];
    d.B = s.batch; d.d_in = s.d_in; d.d_hid = s.d_hid; d.d_out = s.d_out;
    d.dy_parts = s.dy_parts; d.dy_splits = s.dy_splits;
    d.dx = s.dx; d.dw1_parts = s.dw1_parts; d.dw2_parts = s.dw2_parts; d.db1_parts = s.db1_parts; d.db2_parts = s.db2_parts;
    rc = make_tmap_bf16_2d(&args.tmX[i], s.x, (uint64_t)s.d_in, (uint64_t)s.batch, (uint64_t)s.d_in * 2, 64, TW_BM);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&args.tmH[i], s.h, (uint64_t)s.d_hid, (uint64_t)s.batch, (uint64_t)s.d_hid * 2, 64, TW_BM);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&args.tmW2[i], s.w2, (uint64_t)s.d_out, (uint64_t)s.d_hid, (uint64_t)s.d_out * 2, 64, (uint32_t)s.d_hid);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&args.tmW1[i], s.w1, (uint64_t)s.d_hid, (uint64_t)s.d_in, (uint64_t)s.d_hid * 2, 64, (uint32_t)s.d_in);
    if (rc) return rc;
    const int64_t P = ceil_div(s.batch, TW_BM);
    rc = make_tmap_f32_2d(&args.tmDX[i], s.dx, (uint64_t)s.d_in, (uint64_t)s.batch, (uint64_t)s.d_in * 4, 32, 32);
    if (rc) return rc;
    rc = make_tmap_f32_2d(&args.tmDW1[i], s.dw1_parts, (uint64_t)s.d_hid, (uint64_t)(P * s.d_in), (uint64_t)s.d_hid * 4, 32, 32);
    if (rc) return rc;
    rc = make_tmap_f32_2d(&args.tmDW2[i], s.dw2_parts, (uint64_t)s.d_out, (uint64_t)(P * s.d_hid), (uint64_t)s.d_out * 4, 32, 32);
    if (rc) return rc;
    max_b = std::max<int64_t>(max_b, s.batch);
    smem = std::max(smem, bwdt_layout(s.d_in, s.d_hid, s.d_out).total);
  }
  args.trace = g_tower_trace ? g_tower_trace + 16 * 256 : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  TT_CUDA_OK(cudaFuncSetAttribute(tower_mlp2_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid((unsigned)ceil_div(max_b, TW_BM), (unsigned)n);
  TT_PROF("tower_mlp2_bwd_kernel", st);
  TT_CUDA_OK(launch_pdl(tower_mlp2_bwd_kernel, grid, dim3(TWF_THREADS), (size_t)smem, st, args));
  TT_LAUNCH_OK("tower_mlp2_bwd_kernel");
  return TT_OK;
}
