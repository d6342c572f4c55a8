// K3/K4 (fp32 precision) -- tfrs.tasks.Retrieval in-batch softmax cross-entropy, forward and
// backward, fused so the [nq, nc] logits never reach HBM.
//
// One kernel template, three modes.  A CTA keeps a 64-row tile of the stationary operand in
// shared memory and streams the other operand in 64-row tiles:
//   MODE 0 forward : stationary Q rows; online (max, sum-exp) per row; writes lse, pos.
//   MODE 1 dQ      : stationary Q rows; recomputes S, dS = w (softmax - eye) / T, dQ += dS C.
//   MODE 2 dC      : stationary C rows; recomputes S^T, dC += dS^T Q.
// Transforms follow upstream order: /temperature, -log q correction, accidental-hit mask.
// The tcgen05 bf16 version of the same three passes is retrieval_tc.cu.
#include "simt_tile.cuh"

namespace tt {

int tc_retrieval_fwd(const void* q, const void* c, int64_t nq, int64_t nc, int64_t d, float inv_temp,
                     int64_t label_offset, const float* w, const float* logq, const int64_t* cand_ids,
                     float* row_lse, float* row_pos, float* loss, void* ws, int64_t ws_bytes, cudaStream_t st);
int tc_retrieval_bwd(const void* q, const void* c, int64_t nq, int64_t nc,
                     int64_t d, float inv_temp, int64_t label_offset, const float* w, const float* logq,
                     const int64_t* cand_ids, const float* row_lse, float grad_scale, float* dq, float* dc,
                     uint16_t* dq_bf16, uint16_t* dc_bf16, void* ws, int64_t ws_bytes, cudaStream_t st);
int64_t tc_retrieval_workspace_bytes(int64_t nq, int64_t nc, int64_t d);
int64_t tc_retrieval_sync_bytes(int64_t nq, int64_t nc, int64_t d);
void tc_retrieval_bwd_num_splits(int64_t nq, int64_t nc, int64_t d, int* sq, int* sc);
int tc_retrieval_bwd_parts(const void* q, const void* c, int64_t nq, int64_t nc, int64_t d, float inv_temp,
                           int64_t label_offset, const float* w, const float* logq, const int64_t* cand_ids,
                           const float* row_lse, float grad_scale, float* part_q, float* part_c, cudaStream_t st);
int tc_combine_parts(const float* parts, int splits, int64_t rows, int64_t d, float* out_f32, uint16_t* out_bf16, cudaStream_t st);
int64_t tc_retrieval_fwd_dq_workspace_bytes(int64_t nq, int64_t nc, int64_t d);
int tc_retrieval_fwd_dq(const void* q, const void* c, int64_t nq, int64_t nc, int64_t d, float inv_temp,
                        int64_t label_offset, const float* w, float* row_lse, float* row_pos, float* loss, float* dq,
                        void* ws, int64_t ws_bytes, cudaStream_t st, cudaStream_t fin_st);
int tc_retrieval_bwd_dc_scatter(const void* q, const void* c, int64_t nq, int64_t nc, int64_t d, float inv_temp,
                                int64_t label_offset, const float* w, const float* row_lse, float grad_scale,
                                const void* peer_maps, int world, int rank, float* local_part, cudaStream_t st);
int make_tmap_f32_2d_raw(void* out128, uint64_t base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                         uint32_t box_inner, uint32_t box_outer);
int tc_retrieval_bwd_dc_fused(const void* q, const void* c, int64_t nq, int64_t nc, int64_t d, float inv_temp,
                              int64_t label_offset, const float* w, const void* fwd_ws, float grad_scale, float* part_c,
                              cudaStream_t st);

struct RetrievalArgs {
  const float* q; const float* c;
  int64_t nq, nc; int d;
  float inv_temp; int64_t label_offset;
  const float* w; const float* logq; const int64_t* cand_ids;
  const float* lse; float grad_scale;
  float* row_lse; float* row_pos;
  float* dout;
};

__device__ __forceinline__ float transform_logit(const RetrievalArgs& a, float dot, int64_t ci, int64_t label,
                                                 int64_t pos_id) {
  float s = dot * a.inv_temp;
  if (a.logq) s -= __ldg(a.logq + ci);
  if (a.cand_ids && ci != label && __ldg(a.cand_ids + ci) == pos_id) s += TT_MIN_FLOAT;
  return s;
}

template <int MODE, int DJ>
__global__ void __launch_bounds__(256)
retrieval_kernel(const RetrievalArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int d = a.d;
  float* Xs_T = smem;
  float* Ys_T = Xs_T + d * TLD;
  float* Ys = Ys_T + d * TLD;           // MODE > 0 only
  float* dSs_T = Ys + TS * (d + 4);     // MODE > 0 only
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

  const float* X = MODE == 2 ? a.c : a.q;
  const float* Y = MODE == 2 ? a.q : a.c;
  const int64_t nX = MODE == 2 ? a.nc : a.nq;
  const int64_t nY = MODE == 2 ? a.nq : a.nc;
  const int64_t x0 = (int64_t)blockIdx.x * TS;

  load_tile(X, x0, nX, d, Xs_T, nullptr);

  // per-row (stationary) state
  int64_t xi[4];
  float m_run[4], l_run[4];          // MODE 0
  float lse_r[4], w_r[4];            // MODE 1
  int64_t posid_r[4];                // MODE 0/1 (rows are queries)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    xi[i] = x0 + ty * 4 + i;
    m_run[i] = -INFINITY; l_run[i] = 0.f;
    lse_r[i] = 0.f; w_r[i] = 0.f; posid_r[i] = -1;
    if (MODE != 2 && xi[i] < a.nq) {
      if (a.cand_ids) posid_r[i] = __ldg(a.cand_ids + a.label_offset + xi[i]);
      if (MODE == 1) {
        lse_r[i] = __ldg(a.lse + xi[i]);
        w_r[i] = (a.w ? __ldg(a.w + xi[i]) : 1.f) * a.inv_temp * a.grad_scale;
      }
    }
  }
  float acc2[4][DJ][4];
  if (MODE > 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jj = 0; jj < DJ; ++jj)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc2[i][jj][q] = 0.f;
  }

  for (int64_t y0 = 0; y0 < nY; y0 += TS) {
    __syncthreads();                                   // previous tile fully consumed
    load_tile(Y, y0, nY, d, Ys_T, MODE > 0 ? Ys : nullptr);
    __syncthreads();
    float acc[4][4];
    tile_dot(Xs_T, Ys_T, d, tx, ty, acc);

    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (xi[i] >= a.nq) continue;
        const int64_t label = a.label_offset + xi[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t ci = y0 + tx * 4 + j;
          if (ci >= a.nc) continue;
          const float s = transform_logit(a, acc[i][j], ci, label, posid_r[i]);
          if (ci == label) a.row_pos[xi[i]] = s;
          if (s > m_run[i]) { l_run[i] = l_run[i] * expf(m_run[i] - s) + 1.f; m_run[i] = s; }
          else l_run[i] += expf(s - m_run[i]);
        }
      }
    } else {
      // per-column (streamed) state for MODE 2
      float lse_c[4], w_c[4];
      int64_t posid_c[4];
      if (MODE == 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t qi = y0 + tx * 4 + j;
          lse_c[j] = 0.f; w_c[j] = 0.f; posid_c[j] = -1;
          if (qi < a.nq) {
            lse_c[j] = __ldg(a.lse + qi);
            w_c[j] = (a.w ? __ldg(a.w + qi) : 1.f) * a.inv_temp * a.grad_scale;
            if (a.cand_ids) posid_c[j] = __ldg(a.cand_ids + a.label_offset + qi);
          }
        }
      }
      float ds[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t qi = MODE == 1 ? xi[i] : y0 + tx * 4 + j;
          const int64_t ci = MODE == 1 ? y0 + tx * 4 + j : xi[i];
          float v = 0.f;
          if (qi < a.nq && ci < a.nc) {
            const int64_t label = a.label_offset + qi;
            const float s = transform_logit(a, acc[i][j], ci, label, MODE == 1 ? posid_r[i] : posid_c[j]);
            const float p = expf(s - (MODE == 1 ? lse_r[i] : lse_c[j]));
            v = (p - (ci == label ? 1.f : 0.f)) * (MODE == 1 ? w_r[i] : w_c[j]);
          }
          ds[i][j] = v;
        }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(dSs_T + (tx * 4 + j) * TLD + ty * 4) = make_float4(ds[0][j], ds[1][j], ds[2][j], ds[3][j]);
      __syncthreads();
#pragma unroll 2
      for (int c = 0; c < TS; ++c) {
        const float4 dv = *reinterpret_cast<const float4*>(dSs_T + c * TLD + ty * 4);
        const float dvv[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
        for (int jj = 0; jj < DJ; ++jj) {
          const int kq = tx + 16 * jj;
          if (kq * 4 < d) {
            const float4 yv = *reinterpret_cast<const float4*>(Ys + c * (d + 4) + kq * 4);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              acc2[i][jj][0] = fmaf(dvv[i], yv.x, acc2[i][jj][0]);
              acc2[i][jj][1] = fmaf(dvv[i], yv.y, acc2[i][jj][1]);
              acc2[i][jj][2] = fmaf(dvv[i], yv.z, acc2[i][jj][2]);
              acc2[i][jj][3] = fmaf(dvv[i], yv.w, acc2[i][jj][3]);
            }
          }
        }
      }
    }
  }

  if (MODE == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float m = m_run[i], l = l_run[i];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), l2 = __shfl_xor_sync(0xffffffffu, l, o);
        const float mn = fmaxf(m, m2);
        const float e1 = (m == -INFINITY) ? 0.f : expf(m - mn), e2 = (m2 == -INFINITY) ? 0.f : expf(m2 - mn);
        l = l * e1 + l2 * e2;
        m = mn;
      }
      if (tx == 0 && xi[i] < a.nq) a.row_lse[xi[i]] = m + logf(l);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (xi[i] >= nX) continue;
#pragma unroll
      for (int jj = 0; jj < DJ; ++jj) {
        const int kq = tx + 16 * jj;
        if (kq * 4 < d)
          *reinterpret_cast<float4*>(a.dout + xi[i] * d + kq * 4) =
              make_float4(acc2[i][jj][0], acc2[i][jj][1], acc2[i][jj][2], acc2[i][jj][3]);
      }
    }
  }
}

// loss = sum_i w_i (lse_i - pos_i): fixed summation order (strided partials + tree).
__global__ void __launch_bounds__(1024)
loss_reduce_kernel(const float* __restrict__ lse, const float* __restrict__ pos, const float* __restrict__ w,
                   int64_t n, float* __restrict__ loss) {
  __shared__ float part[1024];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 1024) s += (w ? w[i] : 1.f) * (lse[i] - pos[i]);
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = part[0];
}

int launch_loss_reduce(const float* lse, const float* pos, const float* w, int64_t n, float* loss, cudaStream_t st) {
  TT_PROF("loss_reduce_kernel", st);
  loss_reduce_kernel<<<1, 1024, 0, st>>>(lse, pos, w, n, loss);
  TT_LAUNCH_OK("loss_reduce_kernel");
  return TT_OK;
}

template <int MODE>
static int launch_retrieval(const RetrievalArgs& a, cudaStream_t st) {
  const int d = a.d;
  const int64_t nX = MODE == 2 ? a.nc : a.nq;
  size_t smem = (size_t)2 * d * TLD * 4;
  if (MODE > 0) smem += (size_t)(TS * (d + 4) + TS * TLD) * 4;
  const int dj = (d + 63) / 64;
  dim3 grid((unsigned)ceil_div(nX, TS));
#define TT_RL(DJV)                                                                                       \
  {                                                                                                      \
    TT_CUDA_OK(cudaFuncSetAttribute(retrieval_kernel<MODE, DJV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    TT_PROF("retrieval_kernel", st), retrieval_kernel<MODE, DJV><<<grid, 256, smem, st>>>(a);                                             \
  }
  if (MODE == 0) TT_RL(1)
  else if (dj == 1) TT_RL(1)
  else if (dj == 2) TT_RL(2)
  else if (dj == 3) TT_RL(3)
  else TT_RL(4)
#undef TT_RL
  TT_LAUNCH_OK("retrieval_kernel");
  return TT_OK;
}

}  // namespace tt

using namespace tt;

static int check_retrieval_common(const char* fn, int precision, const void* q, const void* c, int64_t nq,
                                  int64_t nc, int64_t d, int64_t label_offset) {
  TT_REQUIRE(precision == TT_F32 || precision == TT_BF16, "%s: unknown precision %d", fn, precision);
  TT_REQUIRE(q && c, "%s: null embeddings", fn);
  TT_REQUIRE(nq > 0 && nc > 0 && d > 0, "%s: empty problem (nq=%lld nc=%lld d=%lld)", fn, (long long)nq, (long long)nc, (long long)d);
  TT_REQUIRE(label_offset >= 0 && label_offset + nq <= nc, "%s: labels [%lld, %lld) exceed the %lld candidates", fn,
             (long long)label_offset, (long long)(label_offset + nq), (long long)nc);
  TT_REQUIRE(aligned16(q) && aligned16(c), "%s: embeddings must be 16-byte aligned", fn);
  if (precision == TT_F32) TT_REQUIRE(d % 4 == 0 && d <= 256, "%s: fp32 path needs d %% 4 == 0 and d <= 256, got %lld", fn, (long long)d);
  return TT_OK;
}

extern "C" int64_t tt_retrieval_workspace_bytes(int32_t precision, int64_t nq, int64_t nc, int64_t d) {
  if (precision == TT_BF16) return tc_retrieval_workspace_bytes(nq, nc, d);
  return 256;
}

extern "C" int tt_retrieval_loss_fwd(int32_t precision, const void* q, const void* c, int64_t nq, int64_t nc,
                                     int64_t d, float inv_temperature, int64_t label_offset,
                                     const float* sample_weight, const float* cand_log_q,
                                     const int64_t* cand_ids, float* row_lse, float* row_pos, float* loss,
                                     void* workspace, int64_t workspace_bytes, void* stream) {
  int rc = check_retrieval_common("tt_retrieval_loss_fwd", precision, q, c, nq, nc, d, label_offset);
  if (rc) return rc;
  TT_REQUIRE(row_lse && row_pos && loss, "tt_retrieval_loss_fwd: null output");
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == TT_BF16)
    return tc_retrieval_fwd(q, c, nq, nc, d, inv_temperature, label_offset, sample_weight, cand_log_q, cand_ids,
                            row_lse, row_pos, loss, workspace, workspace_bytes, st);
  RetrievalArgs a{(const float*)q, (const float*)c, nq, nc, (int)d, inv_temperature, label_offset, sample_weight,
                  cand_log_q, cand_ids, nullptr, 1.f, row_lse, row_pos, nullptr};
  rc = launch_retrieval<0>(a, st);
  if (rc) return rc;
  return launch_loss_reduce(row_lse, row_pos, sample_weight, nq, loss, st);
}

extern "C" int tt_retrieval_loss_bwd(int32_t precision, const void* q, const void* c, int64_t nq, int64_t nc,
                                     int64_t d, float inv_temperature, int64_t label_offset,
                                     const float* sample_weight, const float* cand_log_q,
                                     const int64_t* cand_ids, const float* row_lse, float grad_scale,
                                     float* dq, float* dc, uint16_t* dq_bf16, uint16_t* dc_bf16,
                                     void* workspace, int64_t workspace_bytes, void* stream) {
  int rc = check_retrieval_common("tt_retrieval_loss_bwd", precision, q, c, nq, nc, d, label_offset);
  if (rc) return rc;
  TT_REQUIRE(row_lse && dq && dc, "tt_retrieval_loss_bwd: null buffer");
  TT_REQUIRE(aligned16(dq) && aligned16(dc), "tt_retrieval_loss_bwd: gradients must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == TT_BF16)
    return tc_retrieval_bwd(q, c, nq, nc, d, inv_temperature, label_offset, sample_weight, cand_log_q,
                            cand_ids, row_lse, grad_scale, dq, dc, dq_bf16, dc_bf16, workspace, workspace_bytes, st);
  TT_REQUIRE(!dq_bf16 && !dc_bf16, "tt_retrieval_loss_bwd: bf16 copies are bf16-path outputs");
  RetrievalArgs a{(const float*)q, (const float*)c, nq, nc, (int)d, inv_temperature, label_offset, sample_weight,
                  cand_log_q, cand_ids, row_lse, grad_scale, nullptr, nullptr, dq};
  rc = launch_retrieval<1>(a, st);
  if (rc) return rc;
  a.dout = dc;
  return launch_retrieval<2>(a, st);
}

extern "C" int tt_retrieval_bwd_num_splits(int32_t precision, int64_t nq, int64_t nc, int64_t d, int32_t* dq_splits,
                                           int32_t* dc_splits) {
  TT_REQUIRE(precision == TT_BF16, "tt_retrieval_bwd_num_splits: the split form exists for TT_BF16 only");
  TT_REQUIRE(dq_splits && dc_splits && nq > 0 && nc > 0 && d > 0, "tt_retrieval_bwd_num_splits: bad arguments");
  int sq, sc;
  tc_retrieval_bwd_num_splits(nq, nc, d, &sq, &sc);
  *dq_splits = sq; *dc_splits = sc;
  return TT_OK;
}

extern "C" int tt_retrieval_loss_bwd_parts(int32_t precision, const void* q, const void* c, int64_t nq, int64_t nc,
                                           int64_t d, float inv_temperature, int64_t label_offset,
                                           const float* sample_weight, const float* cand_log_q,
                                           const int64_t* cand_ids, const float* row_lse, float grad_scale,
                                           float* dq_parts, float* dc_parts, void* stream) {
  TT_REQUIRE(precision == TT_BF16, "tt_retrieval_loss_bwd_parts: the split form exists for TT_BF16 only");
  int rc = check_retrieval_common("tt_retrieval_loss_bwd_parts", precision, q, c, nq, nc, d, label_offset);
  if (rc) return rc;
  TT_REQUIRE(row_lse, "tt_retrieval_loss_bwd_parts: null row_lse");
  return tc_retrieval_bwd_parts(q, c, nq, nc, d, inv_temperature, label_offset, sample_weight, cand_log_q, cand_ids,
                                row_lse, grad_scale, dq_parts, dc_parts, (cudaStream_t)stream);
}

extern "C" int64_t tt_retrieval_fwd_dq_workspace_bytes(int64_t nq, int64_t nc, int64_t d) {
  if (nq <= 0 || nc <= 0 || d <= 0) return 0;
  return tc_retrieval_fwd_dq_workspace_bytes(nq, nc, d);
}

extern "C" int tt_retrieval_loss_fwd_dq(const void* q, const void* c, int64_t nq, int64_t nc, int64_t d,
                                        float inv_temperature, int64_t label_offset, const float* sample_weight,
                                        float* row_lse, float* row_pos, float* loss, float* dq, void* workspace,
                                        int64_t workspace_bytes, void* stream, void* finalize_stream) {
  int rc = check_retrieval_common("tt_retrieval_loss_fwd_dq", TT_BF16, q, c, nq, nc, d, label_offset);
  if (rc) return rc;
  return tc_retrieval_fwd_dq(q, c, nq, nc, d, inv_temperature, label_offset, sample_weight, row_lse, row_pos, loss, dq,
                             workspace, workspace_bytes, (cudaStream_t)stream, (cudaStream_t)finalize_stream);
}

extern "C" int tt_retrieval_loss_bwd_dc_fused(const void* q, const void* c, int64_t nq, int64_t nc, int64_t d,
                                              float inv_temperature, int64_t label_offset, const float* sample_weight,
                                              const void* fwd_dq_workspace, float grad_scale, float* dc_parts, void* stream) {
  int rc = check_retrieval_common("tt_retrieval_loss_bwd_dc_fused", TT_BF16, q, c, nq, nc, d, label_offset);
  if (rc) return rc;
  return tc_retrieval_bwd_dc_fused(q, c, nq, nc, d, inv_temperature, label_offset, sample_weight, fwd_dq_workspace,
                                   grad_scale, dc_parts, (cudaStream_t)stream);
}

extern "C" int tt_peer_make_row_maps(const uint64_t* peer_addrs, int32_t world, int64_t rows, int64_t d, void* out_host) {
  TT_REQUIRE(peer_addrs && out_host && world >= 1 && world <= 16 && rows > 0 && d > 0 && d % 32 == 0,
             "tt_peer_make_row_maps: bad arguments (d must be a multiple of 32)");
  for (int r = 0; r < world; ++r) {
    TT_REQUIRE(peer_addrs[r] && (peer_addrs[r] & 15u) == 0, "tt_peer_make_row_maps: address %d null / unaligned", r);
    int rc = make_tmap_f32_2d_raw((char*)out_host + 128 * r, peer_addrs[r], (uint64_t)d, (uint64_t)rows, (uint64_t)d * 4, 32, 32);
    if (rc) return rc;
  }
  return TT_OK;
}

extern "C" int tt_peer_retrieval_bwd_dc(const void* q, const void* c, int64_t nq, int64_t nc, int64_t d, float inv_temperature,
                                        int64_t label_offset, const float* sample_weight, const float* row_lse,
                                        float grad_scale, const void* peer_maps, int32_t world, int32_t rank,
                                        float* scratch, void* stream) {
  int rc = check_retrieval_common("tt_peer_retrieval_bwd_dc", TT_BF16, q, c, nq, nc, d, label_offset);
  if (rc) return rc;
  return tc_retrieval_bwd_dc_scatter(q, c, nq, nc, d, inv_temperature, label_offset, sample_weight, row_lse, grad_scale,
                                     peer_maps, world, rank, scratch, (cudaStream_t)stream);
}

extern "C" int tt_combine_parts_f32(const float* parts, int32_t num_parts, int64_t rows, int64_t d, float* out_f32,
                                    uint16_t* out_bf16, void* stream) {
  return tc_combine_parts(parts, num_parts, rows, d, out_f32, out_bf16, (cudaStream_t)stream);
}

extern "C" int tt_retrieval_workspace_init(int32_t precision, void* workspace, int64_t workspace_bytes, int64_t nq,
                                           int64_t nc, int64_t d, void* stream) {
  TT_REQUIRE(workspace && nq > 0 && nc > 0 && d > 0, "tt_retrieval_workspace_init: bad arguments");
  if (precision != TT_BF16) return TT_OK;
  if (workspace_bytes < tc_retrieval_workspace_bytes(nq, nc, d))
    return set_error(TT_ERR_WORKSPACE, "tt_retrieval_workspace_init: workspace too small");
  TT_CUDA_OK(cudaMemsetAsync(workspace, 0, (size_t)tc_retrieval_sync_bytes(nq, nc, d), (cudaStream_t)stream));
  return TT_OK;
}
