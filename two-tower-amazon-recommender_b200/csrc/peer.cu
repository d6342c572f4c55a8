// Multi-GPU exchange over NVLink peer memory -- the collectives of the row-sharded training
// step written as plain kernels on a SYMMETRIC workspace (same layout on every rank, every rank maps all
// peers' copies: torch.distributed._symmetric_memory), instead of one NCCL call per exchange:
//
//   peer_push        : this rank's block (candidate embeddings, ids) -> the same slot of EVERY rank's workspace
//                      (the all-gather of the in-batch negatives, written by the producer)
//   peer_barrier     : flag exchange through the workspace (st.release.sys / ld.acquire.sys), ~ one NVLink
//                      round trip; also the first phase of the two kernels below
//   peer_combine_scatter : reduce-scatter, producer side -- the ordered sum of the split partials of dC, each row
//                      written straight into the slot [this rank] of the rank that OWNS the candidate; the owner's
//                      backward tower kernel then folds the `world` slots like any other stack of partials
//   peer_push_rows   : embedding-gradient rows -> slot [this rank * b + j] of the rank that owns table row id_j
//   peer_sum         : out[i] = sum over ranks r (in rank order) of source r's buffer[offset + r * stride + i]
//                      (all-reduce of the dense gradients + loss: every rank adds the same slabs in the same order,
//                      so the replicas stay bit-identical)
//   peer_pull_rows   : owner-side gather of gradient rows (kept for comparison: NVLink READS from SM loads are
//                      round trips and ran at ~200 GB/s here; the producer-side WRITES above are posted)
//
// A step of N ranks costs 4 flag barriers and ~(N-1)/N of the data crossing NVLink once, as writes; a cfg2 step
// at N = 2 spent ~200 us in 8 NCCL collectives (20-35 us each, latency-bound) before this.
//
// Epochs: the k-th use of barrier slot i signals the value k, read from a device-resident counter per slot (so
// captured CUDA graphs replay correctly) that the kernel advances once all of its blocks are through.  Waits are
// `>= k` and every rank uses the slots in the same order, so slots need no reset.
#include "common.cuh"

namespace tt {

TT_TL_DEFINE(set_timeline_peer)      // ids: 9 push, 10 / 14 barrier (slot 0 / other), 11 / 12 sum (with / without barrier), 13 pull

constexpr int kMaxWorld = 16;
constexpr long long kPeerSpinLimit = 1ll << 26;   // ~15-60 s of polling: a lost peer traps instead of hanging the box

struct PeerBarrierArgs {
  unsigned long long* const* flag_bases;   // device array [world]: every rank's flag block (peer-mapped)
  unsigned long long* step;                // this rank's device counters: [0, 8) epoch per slot (start at 1), [8, 16) blocks done
  int world, rank, slot;                   // slot < 0: no barrier
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Every block waits; block 0 also signals.  Flags of slot k live at flags[k * kMaxWorld + source rank].
__device__ __forceinline__ void peer_barrier_phase(const PeerBarrierArgs& b) {
  if (b.slot < 0) return;
  const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long*>(b.step + b.slot);
  if (threadIdx.x < b.world) {
    if (blockIdx.x == 0 && blockIdx.y == 0) {
      __threadfence_system();                                  // everything this GPU wrote before is visible first
      st_release_sys(b.flag_bases[threadIdx.x] + b.slot * kMaxWorld + b.rank, epoch);
    }
    const unsigned long long* mine = b.flag_bases[b.rank] + b.slot * kMaxWorld + threadIdx.x;
    long long spins = 0;
    while (ld_acquire_sys(mine) < epoch) {
      if (++spins > kPeerSpinLimit) {
        printf("libtwotower: peer barrier slot %d timed out waiting for rank %d (epoch %llu)\n", b.slot, (int)threadIdx.x, epoch);
        __trap();
      }
    }
  }
  __syncthreads();
  // the last block through advances the slot's epoch (every block has read it by now)
  if (threadIdx.x == 0) {
    const unsigned long long total = (unsigned long long)gridDim.x * gridDim.y;
    if (atomicAdd(b.step + 8 + b.slot, 1ull) == total - 1) {
      b.step[8 + b.slot] = 0;
      b.step[b.slot] = epoch + 1;
    }
  }
}

__global__ void __launch_bounds__(32) peer_barrier_kernel(const PeerBarrierArgs b) {
  long long* const tl = g_tl;
  tl_mark(tl, b.slot == 0 ? 10 : 14, true);
  peer_barrier_phase(b);
  tl_mark(tl, b.slot == 0 ? 10 : 14, false);
}

// ---- push ---------------------------------------------------------------------------------------------
struct PeerPushArgs {
  unsigned char* const* bases;     // device array [world]: workspace base of every rank
  int world, nseg;
  const uint4* src[4];
  long long dst_offset[4];         // bytes from the workspace base (16-byte aligned)
  long long units[4];              // 16-byte units
};
__global__ void __launch_bounds__(256) peer_push_kernel(const PeerPushArgs a) {
  long long* const tl = g_tl;
  tl_mark(tl, 9, true);
  for (int s = 0; s < a.nseg; ++s) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.units[s]; i += (long long)gridDim.x * blockDim.x) {
      const uint4 v = a.src[s][i];
      for (int r = 0; r < a.world; ++r) reinterpret_cast<uint4*>(a.bases[r] + a.dst_offset[s])[i] = v;
    }
  }
  tl_mark(tl, 9, false);
}

// ---- sum over ranks -----------------------------------------------------------------------------------
struct PeerSumArgs {
  unsigned char* const* bases;
  long long offset;                // bytes from the workspace base of the first summed element (16-byte aligned)
  long long stride;                // extra bytes per source index r (slabs inside one workspace), else 0
  long long n4;                    // float4 units
  float4* out;
  PeerBarrierArgs bar;
};
__global__ void __launch_bounds__(256) peer_sum_kernel(const PeerSumArgs a) {
  long long* const tl = g_tl;
  const int tl_id = a.bar.slot >= 0 ? 11 : 12;
  tl_mark(tl, tl_id, true);
  peer_barrier_phase(a.bar);
  const int world = a.bar.world;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v[kMaxWorld];
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r)
      if (r < world) v[r] = __ldcg(reinterpret_cast<const float4*>(a.bases[r] + a.offset + r * a.stride) + i);   // all loads in flight
    float4 s = v[0];
#pragma unroll
    for (int r = 1; r < kMaxWorld; ++r)
      if (r < world) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
    a.out[i] = s;
  }
  tl_mark(tl, tl_id, false);
}

// ---- owner-side gather of gradient rows ---------------------------------------------------------------
struct PeerPullArgs {
  unsigned char* const* bases;
  int ntab, d;                     // d floats per row (multiple of 4)
  long long rows_per_rank;         // b: global position j was produced by rank j / b as its row j % b
  long long n;                     // world * b
  const long long* ids[4];         // GLOBAL ids of the global batch [n]
  long long src_offset[4];         // bytes from the workspace base of the producing rank's [b, d] gradient rows
  float* out[4];                   // [n, d] local; only the rows this rank owns are written
  PeerBarrierArgs bar;
};
__global__ void __launch_bounds__(256) peer_pull_rows_kernel(const PeerPullArgs a) {
  long long* const tl = g_tl;
  tl_mark(tl, 13, true);
  peer_barrier_phase(a.bar);
  const int lane = threadIdx.x & 31;
  const int world = a.bar.world, me = a.bar.rank;
  const int t = blockIdx.y;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long j = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < a.n; j += warps) {
    const long long id = a.ids[t][j];
    if (id < 0 || id % world != me) continue;
    const long long r = j / a.rows_per_rank, lr = j % a.rows_per_rank;
    const float4* src = reinterpret_cast<const float4*>(a.bases[r] + a.src_offset[t]) + lr * (a.d / 4);
    float4* dst = reinterpret_cast<float4*>(a.out[t]) + j * (a.d / 4);
    for (int c = lane; c < a.d / 4; c += 32) dst[c] = __ldcg(src + c);
  }
  tl_mark(tl, 13, false);
}

// ---- producer-side reduce-scatter: ordered sum of split partials, row i -> slot [rank] of owner i / rows_per_rank
struct PeerCombineArgs {
  unsigned char* const* bases;
  const float* parts;              // [splits, rows, d]
  int splits, d, rank;
  long long rows, rows_per_rank;
  long long dst_offset;            // bytes: the owner's [world, rows_per_rank, d] fp32 receive slots
};
__global__ void __launch_bounds__(256) peer_combine_scatter_kernel(const PeerCombineArgs a) {
  long long* const tl = g_tl;
  tl_mark(tl, 7, true);
  const int lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= a.rows) return;
  const long long owner = i / a.rows_per_rank, lr = i % a.rows_per_rank;
  float4* dst = reinterpret_cast<float4*>(a.bases[owner] + a.dst_offset) + ((long long)a.rank * a.rows_per_rank + lr) * (a.d / 4);
  for (int c = lane; c < a.d / 4; c += 32) {
    float4 s = reinterpret_cast<const float4*>(a.parts + i * a.d)[c];
    for (int p = 1; p < a.splits; ++p) {
      const float4 v = reinterpret_cast<const float4*>(a.parts + ((long long)p * a.rows + i) * a.d)[c];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    dst[c] = s;
  }
}

// ---- producer-side routing of embedding-gradient rows to the owners of their table rows
struct PeerPushRowsArgs {
  unsigned char* const* bases;
  int ntab, d, world, rank;
  long long b;
  const long long* ids[4];         // this rank's ids [b]
  const float* src[4];             // this rank's gradient rows [b, d]
  long long dst_offset[4];         // bytes: the owner's [world * b, d] fp32 row buffer of table t
};
__global__ void __launch_bounds__(256) peer_push_rows_kernel(const PeerPushRowsArgs a) {
  long long* const tl = g_tl;
  tl_mark(tl, 13, true);
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.y;
  const long long j = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= a.b) return;
  const long long id = a.ids[t][j];
  if (id < 0) return;
  const float4* src = reinterpret_cast<const float4*>(a.src[t]) + j * (a.d / 4);
  float4* dst = reinterpret_cast<float4*>(a.bases[id % a.world] + a.dst_offset[t]) + ((long long)a.rank * a.b + j) * (a.d / 4);
  for (int c = lane; c < a.d / 4; c += 32) dst[c] = src[c];
}

static int check_bar(const char* fn, const void* flag_bases, const void* step, int world, int rank, int slot) {
  TT_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "%s: bad world / rank (%d / %d)", fn, world, rank);
  TT_REQUIRE(slot < 8, "%s: barrier slot must be < 8", fn);
  TT_REQUIRE(slot < 0 || (flag_bases && step), "%s: null flag table or step counter", fn);
  return TT_OK;
}
static PeerBarrierArgs make_bar(const void* flag_bases, void* step, int world, int rank, int slot) {
  PeerBarrierArgs b;
  b.flag_bases = (unsigned long long* const*)flag_bases;
  b.step = (unsigned long long*)step;
  b.world = world; b.rank = rank; b.slot = slot;
  return b;
}

}  // namespace tt

using namespace tt;

extern "C" int tt_peer_barrier(const void* flag_bases, void* step_counter, int32_t world, int32_t rank, int32_t slot,
                               void* stream) {
  int rc = check_bar("tt_peer_barrier", flag_bases, step_counter, world, rank, slot);
  if (rc) return rc;
  TT_REQUIRE(slot >= 0, "tt_peer_barrier: slot must be >= 0");
  TT_PROF("peer_barrier_kernel", (cudaStream_t)stream);
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(make_bar(flag_bases, step_counter, world, rank, slot));
  TT_LAUNCH_OK("peer_barrier_kernel");
  return TT_OK;
}

extern "C" int tt_peer_push(const void* bases, int32_t world, int32_t nseg, const void* const* src,
                            const int64_t* dst_offset, const int64_t* bytes, void* stream) {
  TT_REQUIRE(bases && world >= 1 && world <= kMaxWorld, "tt_peer_push: bad base table / world");
  TT_REQUIRE(nseg >= 1 && nseg <= 4 && src && dst_offset && bytes, "tt_peer_push: 1..4 segments");
  PeerPushArgs a{};
  a.bases = (unsigned char* const*)bases; a.world = world; a.nseg = nseg;
  long long most = 0;
  for (int s = 0; s < nseg; ++s) {
    TT_REQUIRE(src[s] && aligned16(src[s]) && dst_offset[s] % 16 == 0 && bytes[s] % 16 == 0 && bytes[s] >= 0,
               "tt_peer_push: segment %d must be 16-byte aligned and sized", s);
    a.src[s] = (const uint4*)src[s]; a.dst_offset[s] = dst_offset[s]; a.units[s] = bytes[s] / 16;
    most = std::max<long long>(most, a.units[s]);
  }
  if (most == 0) return TT_OK;
  const unsigned blocks = (unsigned)std::min<long long>(ceil_div(most, 256), (long long)num_sms() * 4);
  TT_PROF("peer_push_kernel", (cudaStream_t)stream);
  peer_push_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
  TT_LAUNCH_OK("peer_push_kernel");
  return TT_OK;
}

extern "C" int tt_peer_sum_f32(const void* bases, int64_t offset_bytes, int64_t stride_bytes, int64_t n, float* out,
                               const void* flag_bases, void* step_counter, int32_t world, int32_t rank, int32_t slot,
                               void* stream) {
  int rc = check_bar("tt_peer_sum_f32", flag_bases, step_counter, world, rank, slot);
  if (rc) return rc;
  TT_REQUIRE(bases && out && aligned16(out) && offset_bytes % 16 == 0 && stride_bytes % 16 == 0 && n >= 0 && n % 4 == 0,
             "tt_peer_sum_f32: null / unaligned buffers or n %% 4 != 0");
  if (n == 0 && slot < 0) return TT_OK;
  PeerSumArgs a{};
  a.bases = (unsigned char* const*)bases; a.offset = offset_bytes; a.stride = stride_bytes; a.n4 = n / 4; a.out = (float4*)out;
  a.bar = make_bar(flag_bases, step_counter, world, rank, slot);
  // one float4 per source per thread where possible: the remote loads are latency-bound (~2-3 us per NVLink round
  // trip), so parallelism, not iteration, hides them
  const unsigned blocks = (unsigned)std::max<long long>(1, std::min<long long>(ceil_div(a.n4, 256), (long long)num_sms() * 32));
  TT_PROF("peer_sum_kernel", (cudaStream_t)stream);
  peer_sum_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
  TT_LAUNCH_OK("peer_sum_kernel");
  return TT_OK;
}

extern "C" int tt_peer_pull_rows(const void* bases, int32_t ntab, const int64_t* const* ids, const int64_t* src_offset,
                                 float* const* out, int64_t rows_per_rank, int64_t d, const void* flag_bases,
                                 void* step_counter, int32_t world, int32_t rank, int32_t slot, void* stream) {
  int rc = check_bar("tt_peer_pull_rows", flag_bases, step_counter, world, rank, slot);
  if (rc) return rc;
  TT_REQUIRE(bases && ntab >= 1 && ntab <= 4 && ids && src_offset && out && rows_per_rank > 0 && d > 0 && d % 4 == 0,
             "tt_peer_pull_rows: bad arguments");
  PeerPullArgs a{};
  a.bases = (unsigned char* const*)bases; a.ntab = ntab; a.d = (int)d; a.rows_per_rank = rows_per_rank;
  a.n = rows_per_rank * world;
  for (int t = 0; t < ntab; ++t) {
    TT_REQUIRE(ids[t] && out[t] && aligned16(out[t]) && src_offset[t] % 16 == 0, "tt_peer_pull_rows: table %d null / unaligned", t);
    a.ids[t] = (const long long*)ids[t]; a.src_offset[t] = src_offset[t]; a.out[t] = out[t];
  }
  a.bar = make_bar(flag_bases, step_counter, world, rank, slot);
  const unsigned bx = (unsigned)std::min<long long>(ceil_div(a.n, 8), (long long)num_sms() * 32);    // ~one row per warp
  TT_PROF("peer_pull_rows_kernel", (cudaStream_t)stream);
  peer_pull_rows_kernel<<<dim3(bx, (unsigned)ntab), 256, 0, (cudaStream_t)stream>>>(a);
  TT_LAUNCH_OK("peer_pull_rows_kernel");
  return TT_OK;
}

extern "C" int tt_peer_combine_scatter(const void* bases, const float* parts, int32_t splits, int64_t rows, int64_t d,
                                       int64_t rows_per_rank, int64_t dst_offset, int32_t world, int32_t rank, void* stream) {
  TT_REQUIRE(bases && parts && aligned16(parts) && splits >= 1 && rows >= 0 && d > 0 && d % 4 == 0, "tt_peer_combine_scatter: bad arguments");
  TT_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world && rows_per_rank > 0 && rows == rows_per_rank * world &&
             dst_offset % 16 == 0, "tt_peer_combine_scatter: rows must be world * rows_per_rank");
  if (rows == 0) return TT_OK;
  PeerCombineArgs a{};
  a.bases = (unsigned char* const*)bases; a.parts = parts; a.splits = splits; a.d = (int)d; a.rank = rank;
  a.rows = rows; a.rows_per_rank = rows_per_rank; a.dst_offset = dst_offset;
  TT_PROF("peer_combine_scatter_kernel", (cudaStream_t)stream);
  peer_combine_scatter_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(a);
  TT_LAUNCH_OK("peer_combine_scatter_kernel");
  return TT_OK;
}

extern "C" int tt_peer_push_rows(const void* bases, int32_t ntab, const int64_t* const* ids, const float* const* src,
                                 const int64_t* dst_offset, int64_t b, int64_t d, int32_t world, int32_t rank, void* stream) {
  TT_REQUIRE(bases && ntab >= 1 && ntab <= 4 && ids && src && dst_offset && b >= 0 && d > 0 && d % 4 == 0, "tt_peer_push_rows: bad arguments");
  TT_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "tt_peer_push_rows: bad world / rank");
  if (b == 0) return TT_OK;
  PeerPushRowsArgs a{};
  a.bases = (unsigned char* const*)bases; a.ntab = ntab; a.d = (int)d; a.world = world; a.rank = rank; a.b = b;
  for (int t = 0; t < ntab; ++t) {
    TT_REQUIRE(ids[t] && src[t] && aligned16(src[t]) && dst_offset[t] % 16 == 0, "tt_peer_push_rows: table %d null / unaligned", t);
    a.ids[t] = (const long long*)ids[t]; a.src[t] = src[t]; a.dst_offset[t] = dst_offset[t];
  }
  TT_PROF("peer_push_rows_kernel", (cudaStream_t)stream);
  peer_push_rows_kernel<<<dim3((unsigned)ceil_div(b, 8), (unsigned)ntab), 256, 0, (cudaStream_t)stream>>>(a);
  TT_LAUNCH_OK("peer_push_rows_kernel");
  return TT_OK;
}
