// K1 -- tower input: Embedding gather + multi-hot sum/mean pooling, all features of a tower
// fused into one pass (out[b] = sum_f pool_f(...)).
//
// Restates tf.nn.embedding_lookup / safe_embedding_lookup_sparse(combiner)
// for the id columns of /root/reference/src/data/preprocessor.py:481-489.
//
// Mapping: one warp per output row.  A row of d fp32 is d/4 float4 chunks; lane l owns
// chunks l, l+32, ... so every table-row read is a run of coalesced 128-bit loads (d=128:
// exactly one LDG.128 per lane = one 512 B row per warp instruction).  Bag members are
// fetched 4 rows at a time (4 independent LDG.128 per lane in flight) and added in member
// order, so the fp32 sum is the sequential sum the oracle computes.  HBM-bound; algorithmic
// bytes = nnz*d*4 (rows) + B*d*s_out + nnz*8 (+ (B+1)*8 offsets).
#include "common.cuh"
#include "gather_row.cuh"

namespace tt {

template <int CH>
__global__ void __launch_bounds__(256)
tower_input_kernel(const FeatureParams p, float* __restrict__ out_f32, uint16_t* __restrict__ out_bf16,
                   int64_t B, int64_t d, int* __restrict__ fault) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int nchunks = (int)(d >> 2);

  for (int64_t b = warp0; b < B; b += nwarps) {
    float4 x[CH];
    gather_row<CH>(p, b, d, lane, nchunks, fault, x);

#pragma unroll
    for (int c = 0; c < CH; ++c) {
      int chunk = lane + 32 * c;
      if (chunk < nchunks) {
        if (out_f32) reinterpret_cast<float4*>(out_f32 + b * d)[chunk] = x[c];
        if (out_bf16) {
          uint2 v;
          v.x = pack_bf16x2(x[c].x, x[c].y);
          v.y = pack_bf16x2(x[c].z, x[c].w);
          reinterpret_cast<uint2*>(out_bf16 + b * d)[chunk] = v;
        }
      }
    }
  }
}

int launch_tower_input(const tt_feature* feats, int n, float* out_f32, uint16_t* out_bf16,
                       int64_t B, int64_t d, int* fault, cudaStream_t stream) {
  FeatureParams p;
  p.n = n;
  for (int i = 0; i < n; ++i) p.f[i] = feats[i];
  if (B == 0) return TT_OK;
  const int warps_per_block = 8;
  int64_t blocks = ceil_div(B, warps_per_block);
  int64_t cap = (int64_t)num_sms() * 8;   // 8 resident CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  dim3 grid((unsigned)blocks), block(256);
  const int64_t chunks = d / 4;
  if (chunks <= 32) TT_PROF("tower_input_kernel", stream), tower_input_kernel<1><<<grid, block, 0, stream>>>(p, out_f32, out_bf16, B, d, fault);
  else if (chunks <= 64) TT_PROF("tower_input_kernel", stream), tower_input_kernel<2><<<grid, block, 0, stream>>>(p, out_f32, out_bf16, B, d, fault);
  else if (chunks <= 128) TT_PROF("tower_input_kernel", stream), tower_input_kernel<4><<<grid, block, 0, stream>>>(p, out_f32, out_bf16, B, d, fault);
  else TT_PROF("tower_input_kernel", stream), tower_input_kernel<8><<<grid, block, 0, stream>>>(p, out_f32, out_bf16, B, d, fault);
  TT_LAUNCH_OK("tower_input_kernel");
  return TT_OK;
}

// ---- stable partition of ids by owner (id % world) for row-sharded tables.
// Single CTA scan per 1024-entry block would need a cross-block prefix; n per rank is small
// (<= 64K), so one CTA walks the batch in 1024-entry strips and keeps running per-owner
// cursors -> stable and deterministic.  Two passes: count, then place.
__global__ void __launch_bounds__(1024)
partition_ids_kernel(const int64_t* __restrict__ ids, int64_t n, int world, int64_t capacity,
                     int64_t* __restrict__ send_ids, int64_t* __restrict__ perm,
                     int64_t* __restrict__ counts, int* __restrict__ overflow) {
  __shared__ int s_warp_cnt[32][8];   // per warp, per owner (world <= 8)
  __shared__ int64_t s_base[8];       // running start of each owner's bucket
  __shared__ int64_t s_total[8];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

  // pass 1: totals
  if (tid < 8) s_total[tid] = 0;
  __syncthreads();
  {
    int local[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) local[o] = 0;
    for (int64_t j = tid; j < n; j += blockDim.x) {
      int o = (int)(ids[j] % world);
#pragma unroll
      for (int k = 0; k < 8; ++k) local[k] += (k == o);
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      int v = local[o];
      for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
      if (lane == 0 && v) atomicAdd((unsigned long long*)&s_total[o], (unsigned long long)v);
    }
  }
  __syncthreads();
  if (tid == 0) {
    int64_t run = 0;
    for (int o = 0; o < world; ++o) {
      s_base[o] = capacity > 0 ? o * capacity : run;
      run += s_total[o];
      counts[o] = s_total[o];
      if (capacity > 0 && s_total[o] > capacity && overflow) *overflow = 1;
    }
  }
  __syncthreads();
  if (capacity > 0)      // padded layout: unused slots carry -1
    for (int64_t j = tid; j < capacity * world; j += blockDim.x) send_ids[j] = -1;
  __syncthreads();

  // pass 2: stable placement, strip by strip
  for (int64_t strip = 0; strip < n; strip += blockDim.x) {
    const int64_t j = strip + tid;
    const bool valid = j < n;
    const int64_t id = valid ? ids[j] : 0;
    const int o = valid ? (int)(id % world) : -1;
    int rank_in_warp = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      unsigned m = __ballot_sync(0xffffffffu, o == k);
      if (o == k) rank_in_warp = __popc(m & ((1u << lane) - 1));
      if (lane == 0) s_warp_cnt[w][k] = __popc(m);
    }
    __syncthreads();
    if (valid) {
      int before = 0;
      for (int ww = 0; ww < w; ++ww) before += s_warp_cnt[ww][o];
      int64_t pos = s_base[o] + before + rank_in_warp;
      const bool fits = capacity <= 0 || pos < (int64_t)(o + 1) * capacity;
      perm[j] = fits ? pos : -1;
      if (fits) send_ids[pos] = id / world;
    }
    __syncthreads();
    if (tid < world) {
      int tot = 0;
      for (int ww = 0; ww < 32; ++ww) tot += s_warp_cnt[ww][tid];
      s_base[tid] += tot;
    }
    __syncthreads();
  }
}

// rows are `chunks` 16-byte units wide (fp32 d/4, bf16 d/8); perm[j] < 0 entries are skipped
__global__ void permute_rows_kernel(const uint4* __restrict__ in, const int64_t* __restrict__ perm,
                                    uint4* __restrict__ out, int64_t n, int64_t chunks, int inverse) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int64_t p = perm[row];
  if (p < 0) return;
  const uint4* src = in + (inverse ? p : row) * chunks;
  uint4* dst = out + (inverse ? row : p) * chunks;
  for (int c = lane; c < (int)chunks; c += 32) dst[c] = src[c];
}

}  // namespace tt

using namespace tt;

extern "C" int tt_tower_input_fwd(const tt_feature* host_feats, int32_t num_feats, float* out_f32,
                                  uint16_t* out_bf16, int64_t B, int64_t d, int32_t* id_fault_flag,
                                  void* stream) {
  TT_REQUIRE(host_feats != nullptr && num_feats >= 1 && num_feats <= TT_MAX_FEATURES,
             "tt_tower_input_fwd: num_feats must be in [1, %d]", TT_MAX_FEATURES);
  TT_REQUIRE(out_f32 != nullptr || out_bf16 != nullptr, "tt_tower_input_fwd: no output buffer");
  TT_REQUIRE(B >= 0 && d > 0 && d % 4 == 0 && d <= 4096, "tt_tower_input_fwd: d must be a multiple of 4 in (0, 4096], got %lld", (long long)d);
  TT_REQUIRE(out_f32 == nullptr || aligned16(out_f32), "tt_tower_input_fwd: out_f32 must be 16-byte aligned");
  TT_REQUIRE(out_bf16 == nullptr || (reinterpret_cast<uintptr_t>(out_bf16) & 7u) == 0, "tt_tower_input_fwd: out_bf16 must be 8-byte aligned");
  for (int i = 0; i < num_feats; ++i) {
    TT_REQUIRE(host_feats[i].table != nullptr && (host_feats[i].shard_world >= 2 || aligned16(host_feats[i].table)), "tt_tower_input_fwd: feature %d table null or not 16-byte aligned", i);
    TT_REQUIRE(host_feats[i].values != nullptr || B == 0, "tt_tower_input_fwd: feature %d has no ids", i);
    TT_REQUIRE(host_feats[i].vocab > 0, "tt_tower_input_fwd: feature %d vocab must be positive", i);
    TT_REQUIRE(host_feats[i].mode == TT_POOL_SUM || host_feats[i].mode == TT_POOL_MEAN, "tt_tower_input_fwd: feature %d bad pooling mode", i);
  }
  return launch_tower_input(host_feats, num_feats, out_f32, out_bf16, B, d, id_fault_flag, (cudaStream_t)stream);
}

extern "C" int tt_embedding_gather_f32(const float* table, const int64_t* ids, float* out, int64_t B,
                                       int64_t d, int64_t vocab, void* stream) {
  tt_feature f{table, ids, nullptr, vocab, TT_POOL_SUM, 0 /* shard_world */};
  return tt_tower_input_fwd(&f, 1, out, nullptr, B, d, nullptr, stream);
}

extern "C" int tt_embedding_gather_bf16(const float* table, const int64_t* ids, uint16_t* out, int64_t B,
                                        int64_t d, int64_t vocab, void* stream) {
  tt_feature f{table, ids, nullptr, vocab, TT_POOL_SUM, 0 /* shard_world */};
  return tt_tower_input_fwd(&f, 1, nullptr, out, B, d, nullptr, stream);
}

extern "C" int tt_embedding_bag_fwd(const float* table, const int64_t* values, const int64_t* offsets,
                                    int32_t mode, void* out, int32_t out_dtype, int64_t num_bags,
                                    int64_t d, int64_t vocab, void* stream) {
  TT_REQUIRE(offsets != nullptr, "tt_embedding_bag_fwd: offsets is NULL");
  TT_REQUIRE(out_dtype == TT_F32 || out_dtype == TT_BF16, "tt_embedding_bag_fwd: bad out_dtype %d", out_dtype);
  tt_feature f{table, values, offsets, vocab, mode, 0};
  return tt_tower_input_fwd(&f, 1, out_dtype == TT_F32 ? (float*)out : nullptr,
                            out_dtype == TT_BF16 ? (uint16_t*)out : nullptr, num_bags, d, nullptr, stream);
}

extern "C" int tt_partition_ids(const int64_t* ids, int64_t n, int32_t world, int64_t capacity,
                                int64_t* send_ids, int64_t* perm, int64_t* counts, int32_t* overflow_flag,
                                void* stream) {
  TT_REQUIRE(world >= 1 && world <= 8, "tt_partition_ids: world must be in [1, 8], got %d", world);
  TT_REQUIRE(n >= 0 && capacity >= 0 && (n == 0 || (ids && send_ids && perm)) && counts, "tt_partition_ids: bad arguments");
  TT_PROF("partition_ids_kernel", (cudaStream_t)stream);
  partition_ids_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(ids, n, world, capacity, send_ids, perm, counts, overflow_flag);
  TT_LAUNCH_OK("partition_ids_kernel");
  return TT_OK;
}

extern "C" int tt_permute_rows(const void* in, const int64_t* perm, void* out, int64_t n,
                               int64_t row_bytes, int32_t inverse, void* stream) {
  TT_REQUIRE(row_bytes > 0 && row_bytes % 16 == 0, "tt_permute_rows: row_bytes must be a multiple of 16");
  TT_REQUIRE(n == 0 || (in && perm && out), "tt_permute_rows: null buffer");
  TT_REQUIRE(aligned16(in) && aligned16(out), "tt_permute_rows: buffers must be 16-byte aligned");
  if (n == 0) return TT_OK;
  TT_PROF("permute_rows_kernel", (cudaStream_t)stream);
  permute_rows_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, (cudaStream_t)stream>>>((const uint4*)in, perm, (uint4*)out, n, row_bytes / 16, inverse);
  TT_LAUNCH_OK("permute_rows_kernel");
  return TT_OK;
}
