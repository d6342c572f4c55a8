// Row gather shared by the stand-alone tower-input kernel (gather.cu) and the fused tower
// kernels (tower_tc.cu): x = sum_f pool_f(table_f[bag_f(b)]) for ONE output row, computed by one
// warp.  Lane l owns float4 chunks l, l+32, ... of the row; bag members are fetched 4 rows at a
// time and added in member order (the sequential fp32 sum the oracle computes).
#pragma once
#include "common.cuh"

namespace tt {

struct FeatureParams {
  tt_feature f[TT_MAX_FEATURES];
  int n;
};

__device__ __forceinline__ float4 ldg_row_chunk(const float* table, int64_t id, int64_t d, int chunk) {
  return __ldg(reinterpret_cast<const float4*>(table + id * d) + chunk);
}
// row `id` of a feature's table: one array, or row shards addressed through local / peer-mapped base pointers
__device__ __forceinline__ float4 ldg_feature_row(const tt_feature& ft, int64_t id, int64_t d, int chunk) {
  if (ft.shard_world >= 2) {
    const float* const* shards = reinterpret_cast<const float* const*>(ft.table);
    const int64_t w = ft.shard_world;
    return ldg_row_chunk(shards[id % w], id / w, d, chunk);
  }
  return ldg_row_chunk(ft.table, id, d, chunk);
}

template <int CH>
__device__ __forceinline__ void gather_row(const FeatureParams& p, int64_t b, int64_t d, int lane, int nchunks,
                                           int* __restrict__ fault, float4 (&x)[CH]) {
#pragma unroll
  for (int c = 0; c < CH; ++c) x[c] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int fi = 0; fi < p.n; ++fi) {
    const tt_feature& ft = p.f[fi];
    float4 e[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) e[c] = make_float4(0.f, 0.f, 0.f, 0.f);

    if (ft.offsets == nullptr) {
      int64_t id = __ldg(ft.values + b);
      if (id < 0 || id >= ft.vocab) {      // id == -1 is padding (zero row); anything else is a fault
        if (lane == 0 && fault && id != -1) atomicExch(fault, 1);
      } else {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          int chunk = lane + 32 * c;
          if (chunk < nchunks) e[c] = ldg_feature_row(ft, id, d, chunk);
        }
      }
    } else {
      const int64_t lo = __ldg(ft.offsets + b), hi = __ldg(ft.offsets + b + 1);
      for (int64_t base = lo; base < hi; base += 32) {
        const int cnt = (int)min((int64_t)32, hi - base);
        int64_t my_id = (lane < cnt) ? __ldg(ft.values + base + lane) : 0;
        if (lane < cnt && (my_id < 0 || my_id >= ft.vocab)) {
          if (fault) atomicExch(fault, 1);
          my_id = -1;
        }
        for (int t = 0; t < cnt; t += 4) {
          float4 r[4][CH];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            int64_t id = __shfl_sync(0xffffffffu, my_id, min(t + u, cnt - 1));
            bool ok = (t + u < cnt) && id >= 0;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              int chunk = lane + 32 * c;
              r[u][c] = (ok && chunk < nchunks) ? ldg_feature_row(ft, id, d, chunk)
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (t + u < cnt) {
#pragma unroll
              for (int c = 0; c < CH; ++c) {
                e[c].x = __fadd_rn(e[c].x, r[u][c].x);
                e[c].y = __fadd_rn(e[c].y, r[u][c].y);
                e[c].z = __fadd_rn(e[c].z, r[u][c].z);
                e[c].w = __fadd_rn(e[c].w, r[u][c].w);
              }
            }
          }
        }
      }
      if (ft.mode == TT_POOL_MEAN && hi > lo) {
        const float L = (float)(hi - lo);
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          e[c].x = __fdiv_rn(e[c].x, L);
          e[c].y = __fdiv_rn(e[c].y, L);
          e[c].z = __fdiv_rn(e[c].z, L);
          e[c].w = __fdiv_rn(e[c].w, L);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (fi == 0) {
        x[c] = e[c];
      } else {
        x[c].x = __fadd_rn(x[c].x, e[c].x);
        x[c].y = __fadd_rn(x[c].y, e[c].y);
        x[c].z = __fadd_rn(x[c].z, e[c].z);
        x[c].w = __fadd_rn(x[c].w, e[c].w);
      }
    }
  }

}

}  // namespace tt
