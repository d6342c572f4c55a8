// K2 (bf16 precision) -- tf.keras.layers.Dense forward / dgrad / wgrad on tcgen05 tensor
// cores: D[M,N] = A[M,K] * B[N,K]^T with bf16 operands (both K-major), fp32 accumulation in
// TMEM, and a fused epilogue (bias, ReLU, ReLU-mask, bf16 / transposed-bf16 / fp32 stores,
// split-K partials).
//
// One 128 x BN output tile per CTA, 192 threads:
//   warp 0     TMA producer: cp.async.bulk.tensor 2D loads of [128 x 64] A and [BN x 64] B
//              tiles (128B swizzle) into a 4-stage ring, completion on `full` mbarriers.
//   warp 1     TMEM allocator + MMA issuer: one elected thread issues tcgen05.mma
//              (cta_group::1, kind::f16, M=128, N=BN, K=16), tcgen05.commit releases ring
//              slots (`empty`) and finally signals `tmem_full`.
//   warps 2-5  epilogue: tcgen05.ld 32x32b (thread = one accumulator row, 32 columns per
//              load), bias / activation, vectorised global stores.
// Tensor-pipe bound for large M; at the tower sizes (K = 128..256) the kernel is short and
// the epilogue stores dominate, which is why all copies a later kernel needs are written here.
#include "tc_common.cuh"

namespace tt {

struct GemmEpilogue {
  const float* bias;       // [N] or null
  const uint16_t* mask;    // bf16 [M, N]: output zeroed where mask <= 0 (ReLU gradient) or null
  uint16_t* out_bf16;      // [M, N] or null
  float* out_f32;          // [splits, M, N] or null
  int M, N, K;
  int k_per_split;         // multiple of 64
  int relu;
};

constexpr int GEMM_BM = 128, GEMM_BK = 64, GEMM_STAGES = 4;

template <int BN>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int BAR_OFF = GEMM_STAGES * (A_BYTES + B_BYTES);
  static constexpr int TOTAL = BAR_OFF + (2 * GEMM_STAGES + 1) * 8 + 16 + 1024;   // + alignment slack
};

// A_MN / B_MN: that operand is stored with its M (resp. N) dimension contiguous, i.e. global
// A is [K, M] (resp. B is [K, N]) row-major; otherwise [M, K] (resp. [N, K]).
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmEpilogue ep) {
  using L = GemmSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + GEMM_STAGES * L::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty = full + GEMM_STAGES;
  uint64_t* tmem_full = empty + GEMM_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * GEMM_BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * ep.k_per_split;
  const int k_end = min(ep.K, k_begin + ep.k_per_split);
  const int nkb = (k_end - k_begin + GEMM_BK - 1) / GEMM_BK;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < GEMM_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one_sync()) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % GEMM_STAGES;
        const uint32_t ph = (kb / GEMM_STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], L::A_BYTES + L::B_BYTES);
        const int k = k_begin + kb * GEMM_BK;
        if (A_MN) {
          for (int c = 0; c < GEMM_BM / 64; ++c) tma_load_2d(sA + s * L::A_BYTES + c * 8192, &tmA, &full[s], m0 + 64 * c, k);
        } else {
          tma_load_2d(sA + s * L::A_BYTES, &tmA, &full[s], k, m0);
        }
        if (B_MN) {
          for (int c = 0; c < BN / 64; ++c) tma_load_2d(sB + s * L::B_BYTES + c * 8192, &tmB, &full[s], n0 + 64 * c, k);
        } else {
          tma_load_2d(sB + s * L::B_BYTES, &tmB, &full[s], k, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % GEMM_STAGES;
        const uint32_t ph = (kb / GEMM_STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint64_t da = A_MN ? umma_desc_mn_sw128(smem_u32(sA + s * L::A_BYTES), 8192) : umma_desc_k_sw128(smem_u32(sA + s * L::A_BYTES));
        const uint64_t db = B_MN ? umma_desc_mn_sw128(smem_u32(sB + s * L::B_BYTES), 8192) : umma_desc_k_sw128(smem_u32(sB + s * L::B_BYTES));
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k)
          umma_bf16_ss(tmem_base, da + (A_MN ? 128 : 2) * k, db + (B_MN ? 128 : 2) * k, idesc, (kb | k) != 0);
        umma_commit(&empty[s]);
      }
      umma_commit(tmem_full);
    }
  } else {
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < ep.M;
    if (nkb > 0) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
    }
    float* out_f32 = ep.out_f32 ? ep.out_f32 + (size_t)blockIdx.z * ep.M * ep.N : nullptr;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= ep.N) break;                 // warp-uniform
      uint32_t r[32];
      if (nkb > 0) {
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      const int col0 = n0 + c0;
      if (ep.bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (col0 + j < ep.N) v[j] += __ldg(ep.bias + col0 + j);
      }
      if (ep.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (ep.mask && row_ok) {
        const uint4* mp = reinterpret_cast<const uint4*>(ep.mask + (size_t)row * ep.N + col0);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (col0 + g * 8 < ep.N) {
            const uint4 mv = __ldg(mp + g);
            const uint32_t w[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float lo = __uint_as_float(w[t] << 16), hi = __uint_as_float(w[t] & 0xffff0000u);
              if (!(lo > 0.f)) v[g * 8 + 2 * t] = 0.f;
              if (!(hi > 0.f)) v[g * 8 + 2 * t + 1] = 0.f;
            }
          }
        }
      }
      if (row_ok) {
        if (ep.out_bf16) {
          uint4* op = reinterpret_cast<uint4*>(ep.out_bf16 + (size_t)row * ep.N + col0);
#pragma unroll
          for (int g = 0; g < 4; ++g)
            if (col0 + g * 8 < ep.N)
              op[g] = make_uint4(pack_bf16x2(v[g * 8], v[g * 8 + 1]), pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]),
                                 pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]));
        }
        if (out_f32) {
          float4* op = reinterpret_cast<float4*>(out_f32 + (size_t)row * ep.N + col0);
#pragma unroll
          for (int g = 0; g < 8; ++g)
            if (col0 + g * 4 < ep.N) op[g] = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// out[p][n] = sum over the p-th row slice of a bf16 [M, N] matrix, fp32, fixed order
// (dbias = colsum(dy) for hidden layers).  grid (ceil(N/64), parts); 256 threads = 32 column
// pairs x 8 row groups.
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const uint16_t* __restrict__ X, float* __restrict__ out, int64_t M, int64_t N, int64_t rows_per_part) {
  __shared__ float part[8][64];
  const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int64_t n = (int64_t)blockIdx.x * 64 + 2 * c;
  const int64_t m_lo = (int64_t)blockIdx.y * rows_per_part, m_hi = min(M, m_lo + rows_per_part);
  float s0 = 0.f, s1 = 0.f;
  if (n < N)
    for (int64_t m = m_lo + g; m < m_hi; m += 8) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(X + m * N + n);
      s0 += __uint_as_float(w << 16);
      s1 += __uint_as_float(w & 0xffff0000u);
    }
  part[g][2 * c] = s0; part[g][2 * c + 1] = s1;
  __syncthreads();
  if (g == 0 && n < N) {
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { t0 += part[i][2 * c]; t1 += part[i][2 * c + 1]; }
    out[(int64_t)blockIdx.y * N + n] = t0;
    out[(int64_t)blockIdx.y * N + n + 1] = t1;
  }
}

// ---- host ----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int make_tmap_2d(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, uint64_t inner, uint64_t outer,
                        uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(TT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, dtype, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(TT_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for [%llu x %llu] stride %llu box {%u,%u}", (int)r,
                     (unsigned long long)outer, (unsigned long long)inner, (unsigned long long)row_stride_bytes,
                     box_inner, box_outer);
  return TT_OK;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                      uint32_t box_inner, uint32_t box_outer) {
  return make_tmap_2d(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, inner, outer, row_stride_bytes, box_inner, box_outer);
}

int make_tmap_f32_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                     uint32_t box_inner, uint32_t box_outer) {
  return make_tmap_2d(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, inner, outer, row_stride_bytes, box_inner, box_outer);
}

// fp32 map from a raw (possibly peer-mapped) address into a 128-byte host buffer (tt_peer_make_row_maps)
int make_tmap_f32_2d_raw(void* out128, uint64_t base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                         uint32_t box_inner, uint32_t box_outer) {
  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is 128 bytes");
  CUtensorMap m;
  int rc = make_tmap_f32_2d(&m, reinterpret_cast<const void*>(static_cast<uintptr_t>(base)), inner, outer, row_stride_bytes, box_inner, box_outer);
  if (rc) return rc;
  memcpy(out128, &m, sizeof(m));
  return TT_OK;
}

// D[M,N] = A * B^T-or-B (+ epilogue); splits over K write out_f32[z].
// a_mn == 0: A is [M, K] row-major; a_mn == 1: A is [K, M] row-major (M contiguous).
// b_mn == 0: B is [N, K] row-major; b_mn == 1: B is [K, N] row-major (N contiguous).
int launch_gemm_tc(const void* A, int a_mn, const void* B, int b_mn, int64_t M, int64_t N, int64_t K, int splits,
                   int k_per_split, GemmEpilogue ep, cudaStream_t st) {
  TT_REQUIRE(K % 8 == 0 && N % 8 == 0 && (!a_mn || M % 8 == 0), "bf16 GEMM needs M, N, K multiples of 8 (got M=%lld N=%lld K=%lld)", (long long)M, (long long)N, (long long)K);
  TT_REQUIRE(aligned16(A) && aligned16(B), "bf16 GEMM operands must be 16-byte aligned");
  constexpr int BN = 128;
  CUtensorMap tmA, tmB;
  int rc = a_mn ? make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)M * 2, 64, GEMM_BK)
                : make_tmap_bf16_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)K * 2, GEMM_BK, GEMM_BM);
  if (rc) return rc;
  rc = b_mn ? make_tmap_bf16_2d(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)N * 2, 64, GEMM_BK)
            : make_tmap_bf16_2d(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)K * 2, GEMM_BK, BN);
  if (rc) return rc;
  ep.M = (int)M; ep.N = (int)N; ep.K = (int)K; ep.k_per_split = k_per_split;
  dim3 grid((unsigned)ceil_div(N, BN), (unsigned)ceil_div(M, GEMM_BM), (unsigned)splits);
  const int smem = GemmSmem<BN>::TOTAL;
#define TT_GEMM(AMN, BMN)                                                                                          \
  {                                                                                                                \
    TT_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, AMN, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    TT_PROF("gemm_tc_kernel", st), gemm_tc_kernel<BN, AMN, BMN><<<grid, 192, smem, st>>>(tmA, tmB, ep);            \
  }
  if (a_mn && b_mn) TT_GEMM(true, true) else if (a_mn) TT_GEMM(true, false) else if (b_mn) TT_GEMM(false, true) else TT_GEMM(false, false)
#undef TT_GEMM
  TT_LAUNCH_OK("gemm_tc_kernel");
  return TT_OK;
}

}  // namespace tt

// Test hook (tests/test_gpu_bf16.py): plain bf16 GEMM in every operand-major combination, fp32 out.
extern "C" int tt_debug_gemm_bf16(const void* A, int32_t a_mn, const void* B, int32_t b_mn, int64_t M, int64_t N, int64_t K,
                                  float* out, void* stream) {
  tt::GemmEpilogue ep{};
  ep.out_f32 = out;
  return tt::launch_gemm_tc(A, a_mn, B, b_mn, M, N, K, 1, (int)tt::round_up(K, tt::GEMM_BK), ep, (cudaStream_t)stream);
}

namespace tt {

int tc_dense_fwd(const void* x, const void* kernel, const float* bias, void* y, float* y_f32,
                 int64_t M, int64_t in_dim, int64_t out_dim, int relu, cudaStream_t stream) {
  GemmEpilogue ep{};
  ep.bias = bias; ep.relu = relu;
  ep.out_bf16 = (uint16_t*)y; ep.out_f32 = y_f32;
  // y[M,out] = x[M,in] * kernel[in,out]: B is the Keras-layout kernel itself (N contiguous)
  return launch_gemm_tc(x, 0, kernel, 1, M, out_dim, in_dim, 1, (int)round_up(in_dim, GEMM_BK), ep, stream);
}

static void wgrad_split(int64_t M, int64_t in_dim, int64_t out_dim, int* parts, int* k_per_split) {
  const int64_t tiles = ceil_div(in_dim, GEMM_BM) * ceil_div(out_dim, 128);
  int64_t want = std::max<int64_t>(1, num_sms() / tiles);
  if (want > 32) want = 32;
  int64_t kps = round_up(ceil_div(M, want), GEMM_BK);
  if (kps < 2 * GEMM_BK) kps = 2 * GEMM_BK;
  *k_per_split = (int)kps;
  *parts = (int)ceil_div(M, kps);
}

int tc_dense_bwd_num_parts(int64_t M, int64_t in_dim, int64_t out_dim) {
  int parts, kps;
  wgrad_split(M, in_dim, out_dim, &parts, &kps);
  return parts;
}

int tc_dense_bwd(const void* dy, const void* x, const void* kernel, void* dx, float* dx_f32,
                 float* dkernel_parts, int num_parts, float* dbias_parts, int64_t M, int64_t in_dim,
                 int64_t out_dim, int relu_mask_x, cudaStream_t stream) {
  TT_REQUIRE(M % 8 == 0, "tt_dense_bwd(bf16): batch must be a multiple of 8 (got %lld)", (long long)M);
  TT_REQUIRE(out_dim % 2 == 0, "tt_dense_bwd(bf16): out_dim must be even");
  int parts, kps;
  wgrad_split(M, in_dim, out_dim, &parts, &kps);
  TT_REQUIRE(num_parts == parts, "tt_dense_bwd(bf16): num_parts=%d, expected tt_dense_bwd_num_parts()=%d", num_parts, parts);
  int rc;
  if (dx || dx_f32) {              // dx[M,in] = dy[M,out] * kernel[in,out]^T  (kernel is [N=in, K=out]: K-major)
    GemmEpilogue ep{};
    ep.mask = relu_mask_x ? (const uint16_t*)x : nullptr;
    ep.out_bf16 = (uint16_t*)dx; ep.out_f32 = dx_f32;
    rc = launch_gemm_tc(dy, 0, kernel, 0, M, in_dim, out_dim, 1, (int)round_up(out_dim, GEMM_BK), ep, stream);
    if (rc) return rc;
  }
  {                                // dkernel[in,out] = x[M,in]^T * dy[M,out]: both operands MN-major, split over M
    GemmEpilogue ep{};
    ep.out_f32 = dkernel_parts;
    rc = launch_gemm_tc(x, 1, dy, 1, in_dim, out_dim, M, parts, kps, ep, stream);
    if (rc) return rc;
  }
  if (dbias_parts) {               // [num_parts, out]: same row slices as the wgrad split
    dim3 grid((unsigned)ceil_div(out_dim, 64), (unsigned)parts);
    TT_PROF("colsum_bf16_kernel", stream);
    colsum_bf16_kernel<<<grid, 256, 0, stream>>>((const uint16_t*)dy, dbias_parts, M, out_dim, kps);
    TT_LAUNCH_OK("colsum_bf16_kernel");
  }
  return TT_OK;
}

}  // namespace tt
