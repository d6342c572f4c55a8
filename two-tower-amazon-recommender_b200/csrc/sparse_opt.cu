// K5 -- sparse embedding gradient: dedup + scatter-add fused with the row-wise optimizer;
// dense Adagrad/Adam for the Dense kernels.
//
// Restates Keras optimizer._deduplicate_sparse_grad (tf.unique, first-occurrence order, +
// unsorted_segment_sum) followed by the sparse update_step of Adagrad (exactly row-wise) or
// LazyAdam.
//
// Three launches, no sort:
//   (1) insert    thread per entry j: open-addressing insert of values[j] into a 2x
//                 over-provisioned hash table (atomicCAS), remember its slot, atomicMin the
//                 slot's first-occurrence position.
//   (2) accumulate warp per entry: add grad[bag(j)] (/L for mean pooling) into the fp32
//                 accumulation row of the id's FIRST occurrence with 128-bit vector atomics
//                 (L2-resident: nnz*d*4 bytes).
//   (3) apply     warp per entry; only first occurrences act: read accumulated row + table
//                 row + slot row(s), write table + slot(s), and clear the accumulation row
//                 and the hash slot so the workspace is clean for the next step.
// HBM-bound.  Algorithmic bytes (Adagrad, U unique of nnz): nnz*d*4 (grad) +
// U*d*4*(2 reads + 2 writes) + nnz*8.
// Duplicate rows are summed with atomics: the SET of updated rows is exact, the fp32 sum
// order over duplicates is not fixed (same as TF's GPU unsorted_segment_sum).
#include "common.cuh"
#include "sparse_ws.cuh"
#include <limits.h>
#include <stdlib.h>
#include <algorithm>

namespace tt {

TT_TL_DEFINE(set_timeline_opt)

__global__ void sparse_ws_init_kernel(SparseWs ws, int64_t nnz, int64_t d) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = i; k < ws.cap; k += stride) { ws.keys[k] = kEmpty; ws.first[k] = INT_MAX; ws.cnt[k] = 0; ws.done[k] = 0; }
  for (int64_t k = i; k < nnz * d; k += stride) ws.accum[k] = 0.f;
}

__global__ void __launch_bounds__(256)
sparse_insert_kernel(SparseWs ws, const int64_t* __restrict__ values, const int64_t* __restrict__ offsets,
                     int64_t num_rows, int64_t nnz, int64_t vocab) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnz) return;
  const int64_t id = values[j];
  if (id < 0 || id >= vocab) { ws.hpos[j] = -1; return; }   // out-of-range ids are dropped
  const uint64_t mask = (uint64_t)ws.cap - 1;
  uint64_t h = mix64((uint64_t)id) & mask;
  while (true) {
    unsigned long long prev = atomicCAS(&ws.keys[h], kEmpty, (unsigned long long)id);
    if (prev == kEmpty || prev == (unsigned long long)id) break;
    h = (h + 1) & mask;
  }
  ws.hpos[j] = (int)h;
  atomicMin(&ws.first[h], (int)j);
  if (offsets) {   // bag(j): last b with offsets[b] <= j
    int64_t lo = 0, hi = num_rows;
    while (hi - lo > 1) {
      int64_t mid = (lo + hi) >> 1;
      if (offsets[mid] <= j) lo = mid; else hi = mid;
    }
    ws.bag_of[j] = (int)lo;
  }
}

__global__ void __launch_bounds__(256)
sparse_accumulate_kernel(SparseWs ws, const int64_t* __restrict__ offsets, int mode, int64_t nnz,
                         int64_t d, const float* __restrict__ grad) {
  const int lane = threadIdx.x & 31;
  const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= nnz) return;
  const int h = ws.hpos[j];
  if (h < 0) return;
  const int leader = ws.first[h];
  int64_t row = j;
  float L = 1.f;
  if (offsets) {
    row = ws.bag_of[j];
    if (mode == TT_POOL_MEAN) L = (float)(offsets[row + 1] - offsets[row]);
  }
  const float4* g = reinterpret_cast<const float4*>(grad + row * d);
  float4* a = reinterpret_cast<float4*>(ws.accum + (int64_t)leader * d);
  for (int c = lane; c < (int)(d >> 2); c += 32) {
    float4 v = __ldg(g + c);
    if (L != 1.f) { v.x = __fdiv_rn(v.x, L); v.y = __fdiv_rn(v.y, L); v.z = __fdiv_rn(v.z, L); v.w = __fdiv_rn(v.w, L); }
    atomicAdd(a + c, v);     // sm_90+: 128-bit vector reduction at L2
  }
}

// fence.acq_rel at device scope: what the ticket protocol needs (__threadfence() is the sequentially consistent fence)
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

struct AdagradRule {
  float* acc; float lr, eps;
  // the arithmetic alone, on a chunk already in registers (the fused step keeps the rows of several entries in flight)
  __device__ __forceinline__ void update(float4& w, float4& a, const float4 g) const {
#define TT_ADAGRAD_1(X)                                                        \
    a.X = __fadd_rn(a.X, __fmul_rn(g.X, g.X));                                 \
    w.X = __fsub_rn(w.X, __fdiv_rn(__fmul_rn(lr, g.X), __fsqrt_rn(__fadd_rn(a.X, eps))));
    TT_ADAGRAD_1(x) TT_ADAGRAD_1(y) TT_ADAGRAD_1(z) TT_ADAGRAD_1(w)
#undef TT_ADAGRAD_1
  }
  __device__ __forceinline__ void apply(float* w_row, int64_t row_off, int c, float4 g) const {
    float4* wp = reinterpret_cast<float4*>(w_row) + c;
    float4* ap = reinterpret_cast<float4*>(acc + row_off) + c;
    float4 w = *wp, a = *ap;
    update(w, a, g);
    *wp = w; *ap = a;
  }
};

struct LazyAdamRule {
  float* m; float* v; float alpha, b1, b2, eps;
  __device__ __forceinline__ void update(float4& w, float4& mm, float4& vv, const float4 g) const {
#define TT_ADAM_1(X)                                                                         \
    mm.X = __fadd_rn(mm.X, __fmul_rn(__fsub_rn(g.X, mm.X), 1.f - b1));                       \
    vv.X = __fadd_rn(vv.X, __fmul_rn(__fsub_rn(__fmul_rn(g.X, g.X), vv.X), 1.f - b2));       \
    w.X = __fsub_rn(w.X, __fdiv_rn(__fmul_rn(mm.X, alpha), __fadd_rn(__fsqrt_rn(vv.X), eps)));
    TT_ADAM_1(x) TT_ADAM_1(y) TT_ADAM_1(z) TT_ADAM_1(w)
#undef TT_ADAM_1
  }
  __device__ __forceinline__ void apply(float* w_row, int64_t row_off, int c, float4 g) const {
    float4* wp = reinterpret_cast<float4*>(w_row) + c;
    float4* mp = reinterpret_cast<float4*>(m + row_off) + c;
    float4* vp = reinterpret_cast<float4*>(v + row_off) + c;
    float4 w = *wp, mm = *mp, vv = *vp;
    update(w, mm, vv, g);
    *wp = w; *mp = mm; *vp = vv;
  }
};

template <class Rule>
__global__ void __launch_bounds__(256)
sparse_apply_kernel(SparseWs ws, Rule rule, float* __restrict__ table, int64_t nnz, int64_t d,
                    uint8_t* __restrict__ first_flag) {
  const int lane = threadIdx.x & 31;
  const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= nnz) return;
  const int h = ws.hpos[j];
  const bool is_first = (h >= 0) && (ws.first[h] == (int)j);
  if (first_flag && lane == 0) first_flag[j] = is_first ? 1 : 0;
  if (!is_first) return;
  const int64_t id = (int64_t)ws.keys[h];
  float4* a = reinterpret_cast<float4*>(ws.accum + j * d);
  for (int c = lane; c < (int)(d >> 2); c += 32) {
    float4 g = a[c];
    a[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    rule.apply(table + id * d, id * d, c, g);
  }
  __syncwarp();
  if (lane == 0) { ws.keys[h] = kEmpty; ws.first[h] = INT_MAX; }
}

template <class Rule>
static int run_sparse(const char* name, Rule rule, float* table, int64_t vocab, int64_t d,
                      const int64_t* values, const int64_t* offsets, int mode, int64_t num_rows,
                      int64_t nnz, const float* grad, void* workspace, int64_t workspace_bytes,
                      uint8_t* first_flag, cudaStream_t stream) {
  TT_REQUIRE(table && values && grad && workspace, "%s: null buffer", name);
  TT_REQUIRE(d > 0 && d % 4 == 0, "%s: d must be a multiple of 4, got %lld", name, (long long)d);
  TT_REQUIRE(aligned16(table) && aligned16(grad) && aligned16(workspace), "%s: buffers must be 16-byte aligned", name);
  TT_REQUIRE(nnz >= 0 && nnz < INT_MAX && num_rows >= 0, "%s: bad sizes", name);
  TT_REQUIRE(offsets != nullptr || nnz == num_rows, "%s: without offsets nnz must equal num_rows", name);
  TT_REQUIRE(mode == TT_POOL_SUM || mode == TT_POOL_MEAN, "%s: bad pooling mode", name);
  SparseWs ws;
  if (!ws_carve(nnz, d, workspace, workspace_bytes, &ws))
    return set_error(TT_ERR_WORKSPACE, "%s: workspace too small (%lld < %lld)", name,
                     (long long)workspace_bytes, (long long)ws_layout(nnz, d, nullptr, nullptr));
  if (nnz == 0) return TT_OK;
  TT_PROF("sparse_insert_kernel", stream);
  sparse_insert_kernel<<<(unsigned)ceil_div(nnz, 256), 256, 0, stream>>>(ws, values, offsets, num_rows, nnz, vocab);
  TT_LAUNCH_OK("sparse_insert_kernel");
  TT_PROF("sparse_accumulate_kernel", stream);
  sparse_accumulate_kernel<<<(unsigned)ceil_div(nnz, 8), 256, 0, stream>>>(ws, offsets, mode, nnz, d, grad);
  TT_LAUNCH_OK("sparse_accumulate_kernel");
  TT_PROF("sparse_apply_kernel", stream);
  sparse_apply_kernel<Rule><<<(unsigned)ceil_div(nnz, 8), 256, 0, stream>>>(ws, rule, table, nnz, d, first_flag);
  TT_LAUNCH_OK("sparse_apply_kernel");
  return TT_OK;
}

// ---- dense optimizers (Dense kernels and biases) --------------------------------------
template <bool ADAM>
__global__ void __launch_bounds__(256)
dense_opt_kernel(float* __restrict__ w, float* __restrict__ s0, float* __restrict__ s1,
                 const float* __restrict__ parts, int num_parts, int64_t rows, int64_t cols,
                 float lr_or_alpha, float b1, float b2, float eps, float l2,
                 uint16_t* __restrict__ shadow) {
  const int64_t n = rows * cols;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float g = 0.f;
  for (int p = 0; p < num_parts; ++p) g = __fadd_rn(g, parts[(int64_t)p * n + i]);
  float wv = w[i];
  if (l2 != 0.f) g = __fadd_rn(g, __fmul_rn(2.f * l2, wv));
  if (ADAM) {
    float m = s0[i], v = s1[i];
    m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), 1.f - b1));
    v = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(g, g), v), 1.f - b2));
    wv = __fsub_rn(wv, __fdiv_rn(__fmul_rn(m, lr_or_alpha), __fadd_rn(__fsqrt_rn(v), eps)));
    s0[i] = m; s1[i] = v;
  } else {
    float a = __fadd_rn(s0[i], __fmul_rn(g, g));
    wv = __fsub_rn(wv, __fdiv_rn(__fmul_rn(lr_or_alpha, g), __fsqrt_rn(__fadd_rn(a, eps))));
    s0[i] = a;
  }
  w[i] = wv;
  if (shadow) shadow[i] = float_to_bf16_bits(wv);
}

// ---- multi-variable forms: one launch per stage for all tables / all dense variables ------
struct SparseMultiVar {
  SparseWs ws;
  float* table; float* s0; float* s1;
  const int64_t* values; const int64_t* offsets;
  const float* grad;
  uint8_t* first_flag;
  int64_t num_rows, nnz, vocab, d;
  int mode;
  int shard;     // 0, or (world << 16) | rank: values are global ids, foreign ones are skipped
};
struct SparseMultiArgs { SparseMultiVar v[TT_MAX_SPARSE_VARS]; };

__global__ void __launch_bounds__(256) sparse_insert_multi_kernel(const __grid_constant__ SparseMultiArgs a) {
  const SparseMultiVar& V = a.v[blockIdx.y];
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= V.nnz) return;
  const int64_t id = V.values[j];
  if (id < 0 || id >= V.vocab) { V.ws.hpos[j] = -1; return; }
  const uint64_t mask = (uint64_t)V.ws.cap - 1;
  uint64_t h = mix64((uint64_t)id) & mask;
  while (true) {
    unsigned long long prev = atomicCAS(&V.ws.keys[h], kEmpty, (unsigned long long)id);
    if (prev == kEmpty || prev == (unsigned long long)id) break;
    h = (h + 1) & mask;
  }
  V.ws.hpos[j] = (int)h;
  atomicMin(&V.ws.first[h], (int)j);
  if (V.offsets) {
    int64_t lo = 0, hi = V.num_rows;
    while (hi - lo > 1) {
      int64_t mid = (lo + hi) >> 1;
      if (V.offsets[mid] <= j) lo = mid; else hi = mid;
    }
    V.ws.bag_of[j] = (int)lo;
  }
}

__global__ void __launch_bounds__(256) sparse_accumulate_multi_kernel(const __grid_constant__ SparseMultiArgs a) {
  const SparseMultiVar& V = a.v[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= V.nnz) return;
  const int h = V.ws.hpos[j];
  if (h < 0) return;
  const int leader = V.ws.first[h];
  int64_t row = j;
  float L = 1.f;
  if (V.offsets) {
    row = V.ws.bag_of[j];
    if (V.mode == TT_POOL_MEAN) L = (float)(V.offsets[row + 1] - V.offsets[row]);
  }
  const float4* g = reinterpret_cast<const float4*>(V.grad + row * V.d);
  float4* acc = reinterpret_cast<float4*>(V.ws.accum + (int64_t)leader * V.d);
  for (int c = lane; c < (int)(V.d >> 2); c += 32) {
    float4 v = __ldg(g + c);
    if (L != 1.f) { v.x = __fdiv_rn(v.x, L); v.y = __fdiv_rn(v.y, L); v.z = __fdiv_rn(v.z, L); v.w = __fdiv_rn(v.w, L); }
    atomicAdd(acc + c, v);
  }
}

template <bool ADAM>
__global__ void __launch_bounds__(256)
sparse_apply_multi_kernel(const __grid_constant__ SparseMultiArgs a, float lr_or_alpha, float b1, float b2, float eps) {
  const SparseMultiVar& V = a.v[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= V.nnz) return;
  const int h = V.ws.hpos[j];
  const bool is_first = (h >= 0) && (V.ws.first[h] == (int)j);
  if (V.first_flag && lane == 0) V.first_flag[j] = is_first ? 1 : 0;
  if (!is_first) return;
  const int64_t id = (int64_t)V.ws.keys[h];
  float4* acc = reinterpret_cast<float4*>(V.ws.accum + j * V.d);
  for (int c = lane; c < (int)(V.d >> 2); c += 32) {
    const float4 g = acc[c];
    acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ADAM) {
      LazyAdamRule rule{V.s0, V.s1, lr_or_alpha, b1, b2, eps};
      rule.apply(V.table + id * V.d, id * V.d, c, g);
    } else {
      AdagradRule rule{V.s0, lr_or_alpha, eps};
      rule.apply(V.table + id * V.d, id * V.d, c, g);
    }
  }
  __syncwarp();
  if (lane == 0) { V.ws.keys[h] = kEmpty; V.ws.first[h] = INT_MAX; }
}

static int run_sparse_multi(const char* name, bool adam, const tt_sparse_var* vars, int n, float lr_or_alpha,
                            float b1, float b2, float eps, cudaStream_t stream) {
  TT_REQUIRE(vars && n >= 1 && n <= TT_MAX_SPARSE_VARS, "%s: num_vars must be in [1, %d]", name, TT_MAX_SPARSE_VARS);
  static thread_local SparseMultiArgs args;
  int64_t max_nnz = 0;
  for (int i = 0; i < n; ++i) {
    const tt_sparse_var& s = vars[i];
    TT_REQUIRE(s.table && s.slot0 && (!adam || s.slot1) && s.values && s.grad && s.workspace, "%s: variable %d has a null buffer", name, i);
    TT_REQUIRE(s.d > 0 && s.d % 4 == 0, "%s: d must be a multiple of 4, got %lld", name, (long long)s.d);
    TT_REQUIRE(aligned16(s.table) && aligned16(s.slot0) && aligned16(s.grad) && aligned16(s.workspace), "%s: buffers must be 16-byte aligned", name);
    TT_REQUIRE(s.nnz >= 0 && s.nnz < INT_MAX && s.num_rows >= 0, "%s: bad sizes", name);
    TT_REQUIRE(s.offsets != nullptr || s.nnz == s.num_rows, "%s: without offsets nnz must equal num_rows", name);
    TT_REQUIRE(s.mode == TT_POOL_SUM || s.mode == TT_POOL_MEAN, "%s: bad pooling mode", name);
    SparseMultiVar& v = args.v[i];
    if (!ws_carve(s.nnz, s.d, s.workspace, s.workspace_bytes, &v.ws))
      return set_error(TT_ERR_WORKSPACE, "%s: workspace of variable %d too small", name, i);
    v.table = s.table; v.s0 = s.slot0; v.s1 = s.slot1; v.values = s.values; v.offsets = s.offsets; v.grad = s.grad;
    v.first_flag = s.first_flag; v.num_rows = s.num_rows; v.nnz = s.nnz; v.vocab = s.vocab; v.d = s.d; v.mode = s.mode;
    v.shard = 0;
    TT_REQUIRE(s.shard == 0, "%s: sharded tables go through tt_optimizer_prepare_sparse + tt_*_step", name);
    max_nnz = std::max<int64_t>(max_nnz, s.nnz);
  }
  if (max_nnz == 0) return TT_OK;
  TT_PROF("sparse_insert_kernel", stream);
  sparse_insert_multi_kernel<<<dim3((unsigned)ceil_div(max_nnz, 256), n), 256, 0, stream>>>(args);
  TT_LAUNCH_OK("sparse_insert_multi_kernel");
  TT_PROF("sparse_accumulate_kernel", stream);
  sparse_accumulate_multi_kernel<<<dim3((unsigned)ceil_div(max_nnz, 8), n), 256, 0, stream>>>(args);
  TT_LAUNCH_OK("sparse_accumulate_multi_kernel");
  TT_PROF("sparse_apply_kernel", stream);
  if (adam) sparse_apply_multi_kernel<true><<<dim3((unsigned)ceil_div(max_nnz, 8), n), 256, 0, stream>>>(args, lr_or_alpha, b1, b2, eps);
  else sparse_apply_multi_kernel<false><<<dim3((unsigned)ceil_div(max_nnz, 8), n), 256, 0, stream>>>(args, lr_or_alpha, b1, b2, eps);
  TT_LAUNCH_OK("sparse_apply_multi_kernel");
  return TT_OK;
}

struct DenseMultiVar {
  float* w; float* s0; float* s1; const float* parts; uint16_t* shadow;
  int64_t n; int num_parts; float l2;
};
struct DenseMultiArgs { DenseMultiVar v[TT_MAX_DENSE_VARS]; };

// thread = 4 consecutive weights (n % 4 == 0) or one weight; the partial sums are read as
// num_parts strided, coalesced float4 streams (L2-resident) and added in index order
template <bool ADAM>
__global__ void __launch_bounds__(256)
dense_opt_multi_kernel(const __grid_constant__ DenseMultiArgs a, float lr_or_alpha, float b1, float b2, float eps) {
  const DenseMultiVar& V = a.v[blockIdx.y];
  const bool vec = (V.n & 3) == 0;
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * (vec ? 4 : 1);
  if (i0 >= V.n) return;
  float g[4] = {0.f, 0.f, 0.f, 0.f};
  if (vec) {
    const float4* pp = reinterpret_cast<const float4*>(V.parts + i0);
    const int64_t stride4 = V.n >> 2;
    int p = 0;
    for (; p + 8 <= V.num_parts; p += 8) {
      float4 t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = __ldg(pp + (int64_t)(p + u) * stride4);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        g[0] = __fadd_rn(g[0], t[u].x); g[1] = __fadd_rn(g[1], t[u].y); g[2] = __fadd_rn(g[2], t[u].z); g[3] = __fadd_rn(g[3], t[u].w);
      }
    }
    for (; p < V.num_parts; ++p) {
      const float4 t = __ldg(pp + (int64_t)p * stride4);
      g[0] = __fadd_rn(g[0], t.x); g[1] = __fadd_rn(g[1], t.y); g[2] = __fadd_rn(g[2], t.z); g[3] = __fadd_rn(g[3], t.w);
    }
  } else {
    for (int p = 0; p < V.num_parts; ++p) g[0] = __fadd_rn(g[0], V.parts[(int64_t)p * V.n + i0]);
  }
  const int cnt = vec ? 4 : 1;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (e >= cnt) break;
    const int64_t i = i0 + e;
    float wv = V.w[i], gv = g[e];
    if (V.l2 != 0.f) gv = __fadd_rn(gv, __fmul_rn(2.f * V.l2, wv));
    if (ADAM) {
      float m = V.s0[i], v = V.s1[i];
      m = __fadd_rn(m, __fmul_rn(__fsub_rn(gv, m), 1.f - b1));
      v = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(gv, gv), v), 1.f - b2));
      wv = __fsub_rn(wv, __fdiv_rn(__fmul_rn(m, lr_or_alpha), __fadd_rn(__fsqrt_rn(v), eps)));
      V.s0[i] = m; V.s1[i] = v;
    } else {
      const float acc = __fadd_rn(V.s0[i], __fmul_rn(gv, gv));
      wv = __fsub_rn(wv, __fdiv_rn(__fmul_rn(lr_or_alpha, gv), __fsqrt_rn(__fadd_rn(acc, eps))));
      V.s0[i] = acc;
    }
    V.w[i] = wv;
    if (V.shadow) V.shadow[i] = float_to_bf16_bits(wv);
  }
}

static int run_dense_multi(const char* name, bool adam, const tt_dense_var* vars, int n, float lr_or_alpha,
                           float b1, float b2, float eps, cudaStream_t stream) {
  TT_REQUIRE(vars && n >= 1 && n <= TT_MAX_DENSE_VARS, "%s: num_vars must be in [1, %d]", name, TT_MAX_DENSE_VARS);
  static thread_local DenseMultiArgs args;
  int64_t max_threads = 0;
  for (int i = 0; i < n; ++i) {
    const tt_dense_var& s = vars[i];
    TT_REQUIRE(s.w && s.slot0 && s.grad_parts && (!adam || s.slot1), "%s: variable %d has a null buffer", name, i);
    TT_REQUIRE(s.n > 0 && s.num_parts >= 1, "%s: variable %d has bad sizes", name, i);
    TT_REQUIRE((s.n & 3) != 0 || aligned16(s.grad_parts), "%s: grad_parts of variable %d must be 16-byte aligned", name, i);
    DenseMultiVar& v = args.v[i];
    v.w = s.w; v.s0 = s.slot0; v.s1 = s.slot1; v.parts = s.grad_parts; v.shadow = s.shadow;
    v.n = s.n; v.num_parts = s.num_parts; v.l2 = s.l2;
    max_threads = std::max<int64_t>(max_threads, (s.n & 3) == 0 ? s.n / 4 : s.n);
  }
  dim3 grid((unsigned)ceil_div(max_threads, 256), (unsigned)n);
  TT_PROF("dense_opt_kernel", stream);
  if (adam) dense_opt_multi_kernel<true><<<grid, 256, 0, stream>>>(args, lr_or_alpha, b1, b2, eps);
  else dense_opt_multi_kernel<false><<<grid, 256, 0, stream>>>(args, lr_or_alpha, b1, b2, eps);
  TT_LAUNCH_OK("dense_opt_multi_kernel");
  return TT_OK;
}

// ---- fused optimizer step ---------------------------------------------------------------------
// (1) sparse_prepare_kernel: the hash insert alone; it needs only the ids, so the host enqueues it on a
//     side stream at lookup time and it overlaps the forward pass.  Besides the slot and the first
//     occurrence it counts the occurrences of every key.
// (2) optimizer_step_kernel: ONE launch for all dense variables and all tables.
//     Table entries: a key that occurs once (the common case) is applied straight from its gradient row:
//     no accumulation traffic at all.  Duplicates add their rows into the accumulation row of the first
//     occurrence with 128-bit reductions and take a ticket; the LAST arriver applies the update and
//     cleans the slot ("last block done", per key).  Dense variables: 4 lanes share a float4 of weights
//     and each folds a quarter of the split-K partials (fixed association: deterministic).
__global__ void __launch_bounds__(256) sparse_prepare_kernel(const __grid_constant__ SparseMultiArgs a) {
  tl_mark(g_tl, 6, true);
  const SparseMultiVar& V = a.v[blockIdx.y];
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= V.nnz) return;
  int64_t id = V.values[j];
  if (V.shard && id >= 0) {                     // global id of a row-sharded table: keep what this rank owns
    const int64_t w = V.shard >> 16, me = V.shard & 0xffff;
    id = (id % w == me) ? id / w : -1;
  }
  if (id < 0 || id >= V.vocab) { V.ws.hpos[j] = -1; return; }
  sparse_prepare_entry(V.ws, j, id);
  if (V.offsets) {
    int64_t lo = 0, hi = V.num_rows;
    while (hi - lo > 1) {
      int64_t mid = (lo + hi) >> 1;
      if (V.offsets[mid] <= j) lo = mid; else hi = mid;
    }
    V.ws.bag_of[j] = (int)lo;
  }
}

struct StepArgs {
  DenseMultiArgs dense;
  SparseMultiArgs sparse;
  int sparse_blocks[TT_MAX_SPARSE_VARS + 1];  // prefix sums of the blocks of each table (the grid starts with them)
  int dense_blocks[TT_MAX_DENSE_VARS + 1];    // prefix sums of the blocks of each dense variable (after the table blocks)
  int n_dense, n_sparse;
  int entries_per_block;                      // table entries handled by one block (<= kStepMaxEntries)
};
static constexpr int kStepMaxEntries = 128;

template <bool ADAM>
__device__ __forceinline__ void apply_row(const SparseMultiVar& V, int64_t id, int c, float4 g, float lr_or_alpha,
                                          float b1, float b2, float eps) {
  if (ADAM) {
    LazyAdamRule rule{V.s0, V.s1, lr_or_alpha, b1, b2, eps};
    rule.apply(V.table + id * V.d, id * V.d, c, g);
  } else {
    AdagradRule rule{V.s0, lr_or_alpha, eps};
    rule.apply(V.table + id * V.d, id * V.d, c, g);
  }
}

template <bool ADAM>
__device__ __forceinline__ void optimizer_step_body(const StepArgs& a, float lr_or_alpha, float b1, float b2, float eps) {
  const int lane = threadIdx.x & 31;
  const int bid = blockIdx.x;
  if (bid >= a.dense_blocks[0]) {
    // ---------------- dense variable (behind the table blocks: short blocks on L2-resident partials make the shorter
    // tail): 4 lanes per float4 of weights, each folds partials q, q+4, ...
    int vi = 0;
    while (bid >= a.dense_blocks[vi + 1]) ++vi;
    const DenseMultiVar& V = a.dense.v[vi];
    const bool vec = (V.n & 3) == 0;
    const int64_t item = (int64_t)(bid - a.dense_blocks[vi]) * 64 + (threadIdx.x >> 2);   // 64 items per block
    const int q = threadIdx.x & 3;
    const int64_t i0 = item * (vec ? 4 : 1);
    const bool live = i0 < V.n;
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    if (live) {
      if (vec) {
        const float4* pp = reinterpret_cast<const float4*>(V.parts + i0);
        const int64_t stride4 = V.n >> 2;
        int p = q;
        for (; p + 12 < V.num_parts; p += 16) {
          float4 t[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) t[u] = __ldg(pp + (int64_t)(p + 4 * u) * stride4);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            g[0] = __fadd_rn(g[0], t[u].x); g[1] = __fadd_rn(g[1], t[u].y); g[2] = __fadd_rn(g[2], t[u].z); g[3] = __fadd_rn(g[3], t[u].w);
          }
        }
        for (; p < V.num_parts; p += 4) {
          const float4 t = __ldg(pp + (int64_t)p * stride4);
          g[0] = __fadd_rn(g[0], t.x); g[1] = __fadd_rn(g[1], t.y); g[2] = __fadd_rn(g[2], t.z); g[3] = __fadd_rn(g[3], t.w);
        }
      } else {
        for (int p = q; p < V.num_parts; p += 4) g[0] = __fadd_rn(g[0], V.parts[(int64_t)p * V.n + i0]);
      }
    }
    // (q0 + q1) + (q2 + q3): fixed order
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      g[e] = __fadd_rn(g[e], __shfl_xor_sync(0xffffffffu, g[e], 1));
      g[e] = __fadd_rn(g[e], __shfl_xor_sync(0xffffffffu, g[e], 2));
    }
    if (!live) return;
    // lane q of the group updates element q of the float4 (scalar variables: lane 0 only)
    if (!vec && q != 0) return;
    const int64_t i = i0 + (vec ? q : 0);
    float gv = vec ? (q == 0 ? g[0] : q == 1 ? g[1] : q == 2 ? g[2] : g[3]) : g[0];
    float wv = V.w[i];
    if (V.l2 != 0.f) gv = __fadd_rn(gv, __fmul_rn(2.f * V.l2, wv));
    if (ADAM) {
      float m = V.s0[i], v = V.s1[i];
      m = __fadd_rn(m, __fmul_rn(__fsub_rn(gv, m), 1.f - b1));
      v = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(gv, gv), v), 1.f - b2));
      wv = __fsub_rn(wv, __fdiv_rn(__fmul_rn(m, lr_or_alpha), __fadd_rn(__fsqrt_rn(v), eps)));
      V.s0[i] = m; V.s1[i] = v;
    } else {
      const float acc = __fadd_rn(V.s0[i], __fmul_rn(gv, gv));
      wv = __fsub_rn(wv, __fdiv_rn(__fmul_rn(lr_or_alpha, gv), __fsqrt_rn(__fadd_rn(acc, eps))));
      V.s0[i] = acc;
    }
    V.w[i] = wv;
    if (V.shadow) V.shadow[i] = float_to_bf16_bits(wv);
    return;
  }
  // ---------------- table entries: `entries_per_block` consecutive entries of one table per block, in phases.
  // Phase 1, THREAD per entry: the chain of small dependent loads (slot position -> key, first occurrence, count; bag ->
  //   bag length) runs once per entry with coalesced accesses; its results go to shared memory.
  // Phase 2 (only if the block holds duplicates), WARP per duplicate entry: the gradient row is added into the first
  //   occurrence's accumulation row (128-bit reductions), 4 rows of a warp in flight.
  // Phase 3, THREAD per duplicate entry: release fence, arrival ticket, acquire fence -- once per BLOCK instead of once
  //   per entry (a device-scope fence is a round trip that waits for every reduction the thread has in flight; one pair
  //   per warp and entry was 30 % of the kernel's stall samples at the cfg3 shape).  The last arriver of an id applies it.
  // Phase 4, WARP per applying entry (ids that occur once, last arrivers), U entries of a warp in flight together: with
  //   everything the earlier phases found in shared memory, the 512-byte rows (gradient or accumulated sum, table,
  //   slots) of U entries are requested back to back -- the kernel lives on the bytes it keeps in flight (a
  //   warp-per-entry version that walked the whole chain alone ran at 2.1-2.4 TB/s at the cfg2 / cfg3 shapes).
  int vi = 0;
  while (vi + 1 < a.n_sparse && bid >= a.sparse_blocks[vi + 1]) ++vi;
  const SparseMultiVar& V = a.sparse.v[vi];
  const int E = a.entries_per_block;
  const int64_t j0 = (int64_t)(bid - a.sparse_blocks[vi]) * E;
  __shared__ long long s_id[kStepMaxEntries];
  __shared__ int s_row[kStepMaxEntries], s_h[kStepMaxEntries], s_leader[kStepMaxEntries], s_count[kStepMaxEntries];
  __shared__ float s_L[kStepMaxEntries];
  int has_dup = 0;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const int64_t j = j0 + e;
    int h = -1, leader = 0, count = 0, row = 0;
    long long id = 0;
    float L = 1.f;
    if (j < V.nnz) {
      h = V.ws.hpos[j];
      if (h >= 0) {
        leader = V.ws.first[h];
        count = V.ws.cnt[h];
        id = (long long)V.ws.keys[h];
        row = (int)j;
        if (V.offsets) {
          row = V.ws.bag_of[j];
          if (V.mode == TT_POOL_MEAN) L = (float)(V.offsets[row + 1] - V.offsets[row]);
        }
        has_dup |= count != 1;
      }
      if (V.first_flag) V.first_flag[j] = (h >= 0 && leader == (int)j) ? 1 : 0;
    }
    s_h[e] = h; s_leader[e] = leader; s_count[e] = count; s_row[e] = row; s_id[e] = id; s_L[e] = L;
  }
  has_dup = __syncthreads_or(has_dup);
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int nch = (int)(V.d >> 2);
  if (nch <= 32) {
    // rows of up to 128 floats: one 16-byte chunk per lane
    const bool mine = lane < nch;
    float4* const T4 = reinterpret_cast<float4*>(V.table);
    float4* const S4 = reinterpret_cast<float4*>(V.s0);
    float4* const M4 = reinterpret_cast<float4*>(V.s1);
    float4* const A4 = reinterpret_cast<float4*>(V.ws.accum);
    const float4* const G4 = reinterpret_cast<const float4*>(V.grad);
    if (has_dup) {
      constexpr int UD = 4;
      bool issued = false;
      for (int e0 = warp; e0 < E; e0 += nw * UD) {
        float4 g[UD];
        bool on[UD];
#pragma unroll
        for (int u = 0; u < UD; ++u) {
          const int e = e0 + u * nw;
          on[u] = e < E && s_h[e] >= 0 && s_count[e] != 1 && mine;
          if (on[u]) g[u] = __ldg(G4 + (int64_t)s_row[e] * nch + lane);
        }
#pragma unroll
        for (int u = 0; u < UD; ++u) {
          if (!on[u]) continue;
          const int e = e0 + u * nw;
          const float L = s_L[e];
          if (L != 1.f) { g[u].x = __fdiv_rn(g[u].x, L); g[u].y = __fdiv_rn(g[u].y, L); g[u].z = __fdiv_rn(g[u].z, L); g[u].w = __fdiv_rn(g[u].w, L); }
          atomicAdd(A4 + (int64_t)s_leader[e] * nch + lane, g[u]);
          issued = true;
        }
      }
      // every thread waits for ITS OWN reductions to be performed at device scope (one fence per warp and block, not
      // per entry); the barrier then orders them before the tickets, which other threads take
      if (issued) fence_acq_rel_gpu();
      __syncthreads();
      for (int e = threadIdx.x; e < E; e += blockDim.x) {
        const int h = s_h[e];
        if (h < 0 || s_count[e] == 1) continue;
        fence_acq_rel_gpu();
        const int ticket = atomicAdd(&V.ws.done[h], 1);
        fence_acq_rel_gpu();
        if (ticket != s_count[e] - 1) s_h[e] = -1;       // an earlier arriver: done
      }
      __syncthreads();
    }
    constexpr int U = 2;
    for (int e0 = warp; e0 < E; e0 += nw * U) {
      int h[U], count[U];
      float4 g[U], w4[U], s4[U], t4[U];
      int64_t ro[U], ao[U];            // float4 index of this lane's chunk in the table / slot rows and in the accumulation row
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + u * nw;
        h[u] = e < E ? s_h[e] : -1;
        count[u] = 0;
        g[u] = w4[u] = s4[u] = t4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        ro[u] = ao[u] = 0;
        if (h[u] < 0) continue;
        count[u] = s_count[e];
        ro[u] = (int64_t)s_id[e] * nch + lane;
        ao[u] = (int64_t)s_leader[e] * nch + lane;
        if (mine) {
          // L1-bypassing loads throughout: the rows have no reuse, and the accumulated sums live in L2
          if (count[u] == 1) {
            g[u] = __ldg(G4 + (int64_t)s_row[e] * nch + lane);
            const float L = s_L[e];
            if (L != 1.f) { g[u].x = __fdiv_rn(g[u].x, L); g[u].y = __fdiv_rn(g[u].y, L); g[u].z = __fdiv_rn(g[u].z, L); g[u].w = __fdiv_rn(g[u].w, L); }
          } else {
            g[u] = __ldcg(A4 + ao[u]);
          }
          w4[u] = __ldcg(T4 + ro[u]);
          s4[u] = __ldcg(S4 + ro[u]);
          if (ADAM) t4[u] = __ldcg(M4 + ro[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (h[u] < 0) continue;
        if (mine) {
          if (count[u] != 1) A4[ao[u]] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ADAM) {
            LazyAdamRule rule{V.s0, V.s1, lr_or_alpha, b1, b2, eps};
            rule.update(w4[u], s4[u], t4[u], g[u]);
            T4[ro[u]] = w4[u]; S4[ro[u]] = s4[u]; M4[ro[u]] = t4[u];
          } else {
            AdagradRule rule{V.s0, lr_or_alpha, eps};
            rule.update(w4[u], s4[u], g[u]);
            T4[ro[u]] = w4[u]; S4[ro[u]] = s4[u];
          }
        }
        if (lane == 0) {
          V.ws.keys[h[u]] = kEmpty; V.ws.first[h[u]] = INT_MAX; V.ws.cnt[h[u]] = 0;
          if (count[u] != 1) V.ws.done[h[u]] = 0;
        }
      }
    }
    return;
  }
  // wider rows: chunk loop, one entry at a time
  for (int e = warp; e < E; e += nw) {
    const int h = s_h[e];
    if (h < 0) continue;
    const int leader = s_leader[e], count = s_count[e];
    const int64_t id = s_id[e];
    const float L = s_L[e];
    const float4* gp = reinterpret_cast<const float4*>(V.grad + (int64_t)s_row[e] * V.d);
    if (count == 1) {
      for (int c = lane; c < nch; c += 32) {
        float4 g = __ldg(gp + c);
        if (L != 1.f) { g.x = __fdiv_rn(g.x, L); g.y = __fdiv_rn(g.y, L); g.z = __fdiv_rn(g.z, L); g.w = __fdiv_rn(g.w, L); }
        apply_row<ADAM>(V, id, c, g, lr_or_alpha, b1, b2, eps);
      }
      if (lane == 0) { V.ws.keys[h] = kEmpty; V.ws.first[h] = INT_MAX; V.ws.cnt[h] = 0; }
      continue;
    }
    float4* acc = reinterpret_cast<float4*>(V.ws.accum + (int64_t)leader * V.d);
    for (int c = lane; c < nch; c += 32) {
      float4 g = __ldg(gp + c);
      if (L != 1.f) { g.x = __fdiv_rn(g.x, L); g.y = __fdiv_rn(g.y, L); g.z = __fdiv_rn(g.z, L); g.w = __fdiv_rn(g.w, L); }
      atomicAdd(acc + c, g);
    }
    __threadfence();
    __syncwarp();
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(&V.ws.done[h], 1);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    if (ticket != count - 1) continue;
    __threadfence();
    for (int c = lane; c < nch; c += 32) {
      const float4 g = __ldcg(acc + c);
      acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      apply_row<ADAM>(V, id, c, g, lr_or_alpha, b1, b2, eps);
    }
    if (lane == 0) { V.ws.keys[h] = kEmpty; V.ws.first[h] = INT_MAX; V.ws.cnt[h] = 0; V.ws.done[h] = 0; }
  }
}

// MINB resident blocks per SM: phase 2 holds the rows of U entries in registers (U * 12 data registers and as many
// addresses), so the register budget, not the occupancy, is what the launch bound protects
template <bool ADAM, int MINB>
__global__ void __launch_bounds__(256, MINB)
optimizer_step_kernel(const __grid_constant__ StepArgs a, float lr_or_alpha, const float* __restrict__ alpha_dev,
                      float b1, float b2, float eps) {
  pdl_wait();                       // launched while the backward tower kernel drains
  pdl_launch_dependents();
  long long* const tl = g_tl;
  tl_mark(tl, 5, true);
  if (alpha_dev != nullptr) lr_or_alpha = __ldg(alpha_dev);      // Adam: bias-corrected step size of THIS iteration (graph-replayable)
  optimizer_step_body<ADAM>(a, lr_or_alpha, b1, b2, eps);
  if (tl) {                         // timeline only: the span runs to the EXIT of the last block
    __syncthreads();
    tl_mark(tl, 5, false);
  }
}

// out_i = ordered sum of the split partials of variable i, for all variables in one launch (before the
// data-parallel all-reduce of the dense gradients).  `w` of the descriptor is the output.
__global__ void __launch_bounds__(256) fold_parts_multi_kernel(const __grid_constant__ DenseMultiArgs a) {
  tl_mark(g_tl, 8, true);
  const DenseMultiVar& V = a.v[blockIdx.y];
  const bool vec = (V.n & 3) == 0;
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * (vec ? 4 : 1);
  if (i0 >= V.n) return;
  if (vec) {
    const float4* pp = reinterpret_cast<const float4*>(V.parts + i0);
    const int64_t stride4 = V.n >> 2;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    int p = 0;
    for (; p + 8 <= V.num_parts; p += 8) {
      float4 t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = __ldg(pp + (int64_t)(p + u) * stride4);
#pragma unroll
      for (int u = 0; u < 8; ++u) { g.x = __fadd_rn(g.x, t[u].x); g.y = __fadd_rn(g.y, t[u].y); g.z = __fadd_rn(g.z, t[u].z); g.w = __fadd_rn(g.w, t[u].w); }
    }
    for (; p < V.num_parts; ++p) {
      const float4 t = __ldg(pp + (int64_t)p * stride4);
      g.x = __fadd_rn(g.x, t.x); g.y = __fadd_rn(g.y, t.y); g.z = __fadd_rn(g.z, t.z); g.w = __fadd_rn(g.w, t.w);
    }
    *reinterpret_cast<float4*>(V.w + i0) = g;
  } else {
    float g = 0.f;
    for (int p = 0; p < V.num_parts; ++p) g = __fadd_rn(g, V.parts[(int64_t)p * V.n + i0]);
    V.w[i0] = g;
  }
}

static int fill_sparse(const char* name, bool adam, const tt_sparse_var* vars, int n, SparseMultiArgs* args, int64_t* max_nnz) {
  TT_REQUIRE(n == 0 || vars, "%s: null variable array", name);
  TT_REQUIRE(n >= 0 && n <= TT_MAX_SPARSE_VARS, "%s: at most %d tables per call", name, TT_MAX_SPARSE_VARS);
  *max_nnz = 0;
  for (int i = 0; i < n; ++i) {
    const tt_sparse_var& s = vars[i];
    TT_REQUIRE(s.table && s.values && s.workspace, "%s: table %d has a null buffer", name, i);
    TT_REQUIRE(s.d > 0 && s.d % 4 == 0, "%s: d must be a multiple of 4, got %lld", name, (long long)s.d);
    TT_REQUIRE(aligned16(s.table) && aligned16(s.workspace), "%s: buffers must be 16-byte aligned", name);
    TT_REQUIRE(s.nnz >= 0 && s.nnz < INT_MAX && s.num_rows >= 0, "%s: bad sizes", name);
    TT_REQUIRE(s.offsets != nullptr || s.nnz == s.num_rows, "%s: without offsets nnz must equal num_rows", name);
    TT_REQUIRE(s.mode == TT_POOL_SUM || s.mode == TT_POOL_MEAN, "%s: bad pooling mode", name);
    SparseMultiVar& v = args->v[i];
    if (!ws_carve(s.nnz, s.d, s.workspace, s.workspace_bytes, &v.ws))
      return set_error(TT_ERR_WORKSPACE, "%s: workspace of table %d too small", name, i);
    v.table = s.table; v.s0 = s.slot0; v.s1 = s.slot1; v.values = s.values; v.offsets = s.offsets; v.grad = s.grad;
    v.first_flag = s.first_flag; v.num_rows = s.num_rows; v.nnz = s.nnz; v.vocab = s.vocab; v.d = s.d; v.mode = s.mode;
    v.shard = s.shard;
    TT_REQUIRE(s.shard == 0 || ((s.shard >> 16) >= 1 && (s.shard & 0xffff) < (s.shard >> 16)), "%s: table %d has a bad shard descriptor", name, i);
    *max_nnz = std::max<int64_t>(*max_nnz, s.nnz);
  }
  (void)adam;
  return TT_OK;
}

static int run_prepare(const tt_sparse_var* vars, int n, cudaStream_t stream) {
  static thread_local SparseMultiArgs args;
  int64_t max_nnz;
  int rc = fill_sparse("tt_optimizer_prepare_sparse", false, vars, n, &args, &max_nnz);
  if (rc) return rc;
  if (n == 0 || max_nnz == 0) return TT_OK;
  TT_PROF("sparse_prepare_kernel", stream);
  sparse_prepare_kernel<<<dim3((unsigned)ceil_div(max_nnz, 256), n), 256, 0, stream>>>(args);
  TT_LAUNCH_OK("sparse_prepare_kernel");
  return TT_OK;
}

static int run_step(const char* name, bool adam, const tt_dense_var* dense, int nd, const tt_sparse_var* sparse, int ns,
                    float lr_or_alpha, const float* alpha_dev, float b1, float b2, float eps, cudaStream_t stream) {
  TT_REQUIRE(nd >= 0 && nd <= TT_MAX_DENSE_VARS && (nd == 0 || dense), "%s: at most %d dense variables per call", name, TT_MAX_DENSE_VARS);
  static thread_local StepArgs args;
  int64_t max_nnz;
  int rc = fill_sparse(name, adam, sparse, ns, &args.sparse, &max_nnz);
  if (rc) return rc;
  for (int i = 0; i < ns; ++i)
    TT_REQUIRE(sparse[i].slot0 && sparse[i].grad && aligned16(sparse[i].slot0) && aligned16(sparse[i].grad) && (!adam || sparse[i].slot1),
               "%s: table %d has a null or unaligned slot / gradient", name, i);
  // entries per block: 32, or more when that would be over two waves of table blocks (short blocks keep the tail short;
  // more entries per block amortise phase 1 and the block launch)
  static const int env_epb = [] { const char* e = getenv("TT_OPT_ENTRIES"); return e ? atoi(e) : 0; }();
  int epb = 32;
  if (env_epb == 32 || env_epb == 64 || env_epb == 128) epb = env_epb;
  else {
    int64_t total = 0;
    for (int i = 0; i < ns; ++i) total += sparse[i].nnz;
    while (epb < kStepMaxEntries && ceil_div(total, (int64_t)epb) > 2 * 148 * 4) epb *= 2;
  }
  args.entries_per_block = epb;
  int blocks = 0;
  args.sparse_blocks[0] = 0;
  for (int i = 0; i < ns; ++i) {
    blocks += (int)ceil_div(sparse[i].nnz, (int64_t)epb);
    args.sparse_blocks[i + 1] = blocks;
  }
  for (int i = ns; i < TT_MAX_SPARSE_VARS; ++i) args.sparse_blocks[i + 1] = blocks;
  args.dense_blocks[0] = blocks;
  for (int i = 0; i < nd; ++i) {
    const tt_dense_var& s = dense[i];
    TT_REQUIRE(s.w && s.slot0 && s.grad_parts && (!adam || s.slot1), "%s: dense variable %d has a null buffer", name, i);
    TT_REQUIRE(s.n > 0 && s.num_parts >= 1, "%s: dense variable %d has bad sizes", name, i);
    TT_REQUIRE((s.n & 3) != 0 || aligned16(s.grad_parts), "%s: grad_parts of variable %d must be 16-byte aligned", name, i);
    DenseMultiVar& v = args.dense.v[i];
    v.w = s.w; v.s0 = s.slot0; v.s1 = s.slot1; v.parts = s.grad_parts; v.shadow = s.shadow;
    v.n = s.n; v.num_parts = s.num_parts; v.l2 = s.l2;
    const int64_t items = (s.n & 3) == 0 ? s.n / 4 : s.n;
    blocks += (int)ceil_div(items, 64);
    args.dense_blocks[i + 1] = blocks;
  }
  for (int i = nd; i < TT_MAX_DENSE_VARS; ++i) args.dense_blocks[i + 1] = blocks;
  args.n_dense = nd; args.n_sparse = ns;
  if (blocks == 0) return TT_OK;
  TT_PROF("optimizer_step_kernel", stream);
  static const int minb = [] { const char* e = getenv("TT_OPT_MINB"); return e ? atoi(e) : 4; }();     // A/B runs: 2, 3 or 4
#define TT_STEP_LAUNCH(A, M) TT_CUDA_OK(launch_pdl(optimizer_step_kernel<A, M>, dim3(blocks), dim3(256), (size_t)0, stream, args, lr_or_alpha, alpha_dev, b1, b2, eps))
  if (adam) { if (minb == 2) TT_STEP_LAUNCH(true, 2); else if (minb == 4) TT_STEP_LAUNCH(true, 4); else TT_STEP_LAUNCH(true, 3); }
  else { if (minb == 2) TT_STEP_LAUNCH(false, 2); else if (minb == 4) TT_STEP_LAUNCH(false, 4); else TT_STEP_LAUNCH(false, 3); }
#undef TT_STEP_LAUNCH
  TT_LAUNCH_OK("optimizer_step_kernel");
  return TT_OK;
}

// Keras Adam: alpha_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t) with t = ++iterations.  The counter lives in device memory
// so that a captured CUDA graph advances it on every replay (a host-computed alpha would be frozen at capture time).
__global__ void adam_bias_correction_kernel(long long* step, float lr, float b1, float b2, float* alpha) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const long long t = *step + 1;
    *step = t;
    *alpha = (float)((double)lr * sqrt(1.0 - pow((double)b2, (double)t)) / (1.0 - pow((double)b1, (double)t)));
  }
}

__global__ void __launch_bounds__(1024)
sum_squares_kernel(const float* __restrict__ x, int64_t n, float scale, float* __restrict__ out, int accumulate) {
  __shared__ float part[1024];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 1024) s = fmaf(x[i], x[i], s);
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + scale * part[0];
}

}  // namespace tt

using namespace tt;

extern "C" int tt_sum_squares(const float* x, int64_t n, float scale, float* out, int32_t accumulate, void* stream) {
  TT_REQUIRE(x && out && n >= 0, "tt_sum_squares: bad arguments");
  TT_PROF("sum_squares_kernel", (cudaStream_t)stream);
  sum_squares_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x, n, scale, out, accumulate);
  TT_LAUNCH_OK("sum_squares_kernel");
  return TT_OK;
}

extern "C" int64_t tt_sparse_workspace_bytes(int64_t nnz, int64_t d) {
  if (nnz < 0 || d <= 0) return 0;
  return ws_layout(nnz, d, nullptr, nullptr);
}

extern "C" int tt_sparse_workspace_init(void* workspace, int64_t workspace_bytes, int64_t nnz,
                                        int64_t d, void* stream) {
  TT_REQUIRE(workspace && aligned16(workspace), "tt_sparse_workspace_init: workspace null or unaligned");
  TT_REQUIRE(nnz >= 0 && d > 0, "tt_sparse_workspace_init: bad sizes");
  SparseWs ws;
  if (!ws_carve(nnz, d, workspace, workspace_bytes, &ws))
    return set_error(TT_ERR_WORKSPACE, "tt_sparse_workspace_init: workspace too small");
  // the whole capacity of the buffer is initialised: later batches may hold any entry count up to it
  TT_PROF("sparse_ws_init_kernel", (cudaStream_t)stream);
  sparse_ws_init_kernel<<<num_sms() * 4, 256, 0, (cudaStream_t)stream>>>(ws, ws_capacity(workspace_bytes, d), d);
  TT_LAUNCH_OK("sparse_ws_init_kernel");
  return TT_OK;
}

extern "C" int tt_sparse_adagrad_update(float* table, float* accum, int64_t vocab, int64_t d,
                                        const int64_t* values, const int64_t* offsets, int32_t mode,
                                        int64_t num_rows, int64_t nnz, const float* grad, float lr,
                                        float eps, void* workspace, int64_t workspace_bytes,
                                        uint8_t* first_flag, void* stream) {
  TT_REQUIRE(accum && aligned16(accum), "tt_sparse_adagrad_update: accumulator null or unaligned");
  AdagradRule rule{accum, lr, eps};
  return run_sparse("tt_sparse_adagrad_update", rule, table, vocab, d, values, offsets, mode, num_rows,
                    nnz, grad, workspace, workspace_bytes, first_flag, (cudaStream_t)stream);
}

extern "C" int tt_sparse_lazy_adam_update(float* table, float* m, float* v, int64_t vocab, int64_t d,
                                          const int64_t* values, const int64_t* offsets, int32_t mode,
                                          int64_t num_rows, int64_t nnz, const float* grad, float alpha,
                                          float beta1, float beta2, float eps, void* workspace,
                                          int64_t workspace_bytes, uint8_t* first_flag, void* stream) {
  TT_REQUIRE(m && v && aligned16(m) && aligned16(v), "tt_sparse_lazy_adam_update: slots null or unaligned");
  LazyAdamRule rule{m, v, alpha, beta1, beta2, eps};
  return run_sparse("tt_sparse_lazy_adam_update", rule, table, vocab, d, values, offsets, mode, num_rows,
                    nnz, grad, workspace, workspace_bytes, first_flag, (cudaStream_t)stream);
}

static int dense_opt(bool adam, float* w, float* s0, float* s1, const float* parts, int num_parts,
                     int64_t rows, int64_t cols, float lr, float b1, float b2, float eps, float l2,
                     uint16_t* shadow, cudaStream_t stream) {
  TT_REQUIRE(w && s0 && parts && (!adam || s1), "dense optimizer: null buffer");
  TT_REQUIRE(rows > 0 && cols > 0 && num_parts >= 1, "dense optimizer: bad sizes");
  const int64_t n = rows * cols;
  unsigned blocks = (unsigned)ceil_div(n, 256);
  if (adam) TT_PROF("dense_opt_kernel", stream), dense_opt_kernel<true><<<blocks, 256, 0, stream>>>(w, s0, s1, parts, num_parts, rows, cols, lr, b1, b2, eps, l2, shadow);
  else TT_PROF("dense_opt_kernel", stream), dense_opt_kernel<false><<<blocks, 256, 0, stream>>>(w, s0, s1, parts, num_parts, rows, cols, lr, b1, b2, eps, l2, shadow);
  TT_LAUNCH_OK("dense_opt_kernel");
  return TT_OK;
}

extern "C" int tt_dense_adagrad_update(float* w, float* accum, const float* grad_parts, int32_t num_parts,
                                       int64_t rows, int64_t cols, float lr, float eps, float l2,
                                       uint16_t* shadow, void* stream) {
  return dense_opt(false, w, accum, nullptr, grad_parts, num_parts, rows, cols, lr, 0.f, 0.f, eps, l2,
                   shadow, (cudaStream_t)stream);
}

extern "C" int tt_dense_adam_update(float* w, float* m, float* v, const float* grad_parts, int32_t num_parts,
                                    int64_t rows, int64_t cols, float alpha, float beta1, float beta2,
                                    float eps, float l2, uint16_t* shadow, void* stream) {
  return dense_opt(true, w, m, v, grad_parts, num_parts, rows, cols, alpha, beta1, beta2, eps, l2,
                   shadow, (cudaStream_t)stream);
}

extern "C" int tt_sparse_adagrad_update_multi(const tt_sparse_var* host_vars, int32_t num_vars, float lr, float eps,
                                              void* stream) {
  return run_sparse_multi("tt_sparse_adagrad_update_multi", false, host_vars, num_vars, lr, 0.f, 0.f, eps, (cudaStream_t)stream);
}

extern "C" int tt_sparse_lazy_adam_update_multi(const tt_sparse_var* host_vars, int32_t num_vars, float alpha,
                                                float beta1, float beta2, float eps, void* stream) {
  return run_sparse_multi("tt_sparse_lazy_adam_update_multi", true, host_vars, num_vars, alpha, beta1, beta2, eps, (cudaStream_t)stream);
}

extern "C" int tt_dense_adagrad_update_multi(const tt_dense_var* host_vars, int32_t num_vars, float lr, float eps,
                                             void* stream) {
  return run_dense_multi("tt_dense_adagrad_update_multi", false, host_vars, num_vars, lr, 0.f, 0.f, eps, (cudaStream_t)stream);
}

extern "C" int tt_dense_adam_update_multi(const tt_dense_var* host_vars, int32_t num_vars, float alpha, float beta1,
                                          float beta2, float eps, void* stream) {
  return run_dense_multi("tt_dense_adam_update_multi", true, host_vars, num_vars, alpha, beta1, beta2, eps, (cudaStream_t)stream);
}

extern "C" int tt_optimizer_prepare_sparse(const tt_sparse_var* host_vars, int32_t num_vars, void* stream) {
  return run_prepare(host_vars, num_vars, (cudaStream_t)stream);
}

extern "C" int tt_adagrad_step(const tt_dense_var* host_dense, int32_t num_dense, const tt_sparse_var* host_sparse,
                               int32_t num_sparse, float lr, float eps, void* stream) {
  return run_step("tt_adagrad_step", false, host_dense, num_dense, host_sparse, num_sparse, lr, nullptr, 0.f, 0.f, eps, (cudaStream_t)stream);
}

extern "C" int tt_lazy_adam_step(const tt_dense_var* host_dense, int32_t num_dense, const tt_sparse_var* host_sparse,
                                 int32_t num_sparse, float alpha, const float* alpha_device, float beta1, float beta2, float eps,
                                 void* stream) {
  return run_step("tt_lazy_adam_step", true, host_dense, num_dense, host_sparse, num_sparse, alpha, alpha_device, beta1, beta2, eps, (cudaStream_t)stream);
}

extern "C" int tt_adam_bias_correction(int64_t* step_device, float lr, float beta1, float beta2, float* alpha_device, void* stream) {
  TT_REQUIRE(step_device && alpha_device, "tt_adam_bias_correction: null buffer");
  TT_PROF("adam_bias_correction_kernel", (cudaStream_t)stream);
  adam_bias_correction_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((long long*)step_device, lr, beta1, beta2, alpha_device);
  TT_LAUNCH_OK("adam_bias_correction_kernel");
  return TT_OK;
}

extern "C" int tt_fold_parts_multi(const tt_dense_var* host_vars, int32_t num_vars, void* stream) {
  TT_REQUIRE(host_vars && num_vars >= 1 && num_vars <= TT_MAX_DENSE_VARS, "tt_fold_parts_multi: num_vars must be in [1, %d]", TT_MAX_DENSE_VARS);
  static thread_local DenseMultiArgs args;
  int64_t max_threads = 0;
  for (int i = 0; i < num_vars; ++i) {
    const tt_dense_var& s = host_vars[i];
    TT_REQUIRE(s.w && s.grad_parts && s.n > 0 && s.num_parts >= 1, "tt_fold_parts_multi: variable %d is incomplete", i);
    TT_REQUIRE((s.n & 3) != 0 || (aligned16(s.grad_parts) && aligned16(s.w)), "tt_fold_parts_multi: variable %d must be 16-byte aligned", i);
    DenseMultiVar& v = args.v[i];
    v.w = s.w; v.parts = s.grad_parts; v.n = s.n; v.num_parts = s.num_parts;
    max_threads = std::max<int64_t>(max_threads, (s.n & 3) == 0 ? s.n / 4 : s.n);
  }
  TT_PROF("fold_parts_multi_kernel", (cudaStream_t)stream);
  fold_parts_multi_kernel<<<dim3((unsigned)ceil_div(max_threads, 256), (unsigned)num_vars), 256, 0, (cudaStream_t)stream>>>(args);
  TT_LAUNCH_OK("fold_parts_multi_kernel");
  return TT_OK;
}
