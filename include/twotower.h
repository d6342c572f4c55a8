/*
 * twotower.h -- C-ABI of libtwotower.so: the B200 (sm_100a) two-tower training + retrieval
 * hot path.
 *
 * Boundary.  The reference repository declares this path but ships no code and no FFI for it
 * (/root/reference/src/models/__init__.py:1, src/training/__init__.py:1,
 * src/serving/__init__.py:1, src/evaluation/__init__.py:1), so each entry point below cites
 * the upstream TensorFlow-Recommenders / Keras / FAISS routine that the reference's declared
 * dependencies (/root/reference/pyproject.toml:22,24,39) would dispatch for that step, plus the
 * in-repo line that parameterises it.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - extern "C", plain pointers + sizes.  Every pointer is a DEVICE pointer unless named
 *     host_*.  The caller owns every buffer, including workspaces (sizes from
 *     tt_*_workspace_bytes).  The library allocates nothing persistent.
 *   - Every call enqueues on `stream` (a cudaStream_t passed as void*) and returns without
 *     synchronising; it is CUDA-graph capturable.
 *   - Return value: 0 = TT_OK, negative = error; the message is in tt_last_error()
 *     (thread-local).  Nothing throws or exits across the ABI.  There is no CPU fallback.
 *   - Matrices are row-major.  bf16 values are passed as uint16_t bit patterns.
 */
#ifndef TWOTOWER_H_
#define TWOTOWER_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TT_VERSION 120 /* 0.2.0 */
#define TT_TOPK_RERANK_MARGIN 16

enum tt_status {
  TT_OK = 0,
  TT_ERR_INVALID_ARG = -1,
  TT_ERR_CUDA = -2,
  TT_ERR_UNSUPPORTED = -3,
  TT_ERR_WORKSPACE = -4
};

enum tt_dtype { TT_F32 = 0, TT_BF16 = 1 };
enum tt_pool { TT_POOL_SUM = 0, TT_POOL_MEAN = 1 };

int tt_version(void);
const char* tt_last_error(void);
/* 0 when the current CUDA device is compute capability 10.x, TT_ERR_UNSUPPORTED otherwise. */
int tt_device_check(void);

/* ---------------------------------------------------------------------------------------
 * K1  tower input: Embedding gather and multi-hot sum/mean pooling.
 * Replaces tf.keras.layers.Embedding -> tf.nn.embedding_lookup and
 * safe_embedding_lookup_sparse(combiner); ids are the int64 codes of
 * /root/reference/src/data/preprocessor.py:481-482,485-489.
 * ------------------------------------------------------------------------------------- */
typedef struct tt_feature {
  const float* table;     /* [vocab, d] fp32; or, when shard_world >= 2, a DEVICE array of shard_world
                           * base pointers (const float* const*) of the row shards of the table, local
                           * and peer-mapped (NVLink P2P): row id lives in shard id % shard_world at local
                           * row id / shard_world -- the lookup of a row-sharded table then needs no
                           * all-to-all, the gather kernel loads the rows straight from the owners */
  const int64_t* values;  /* ids: [B] when offsets == NULL, else CSR values [nnz] */
  const int64_t* offsets; /* NULL (one id per row) or CSR offsets [B+1] */
  int64_t vocab;          /* GLOBAL number of rows */
  int32_t mode;           /* tt_pool, used when offsets != NULL */
  int32_t shard_world;    /* 0 / 1: one array */
} tt_feature;

#define TT_MAX_FEATURES 8

/* out[b,:] = sum_f pool_f(table_f[bag_f(b)]).  Either output may be NULL (not both).
 * `host_feats` is a HOST array of num_feats descriptors (copied into kernel parameters).
 * d % 4 == 0, rows 16-byte aligned.  Ids outside [0, vocab) raise a device-side flag that
 * sets *id_fault_flag = 1 (a caller-owned device int, nullable) and read as zero rows
 * (TF raises on CPU; a GPU kernel cannot). */
int tt_tower_input_fwd(const tt_feature* host_feats, int32_t num_feats, float* out_f32,
                       uint16_t* out_bf16, int64_t B, int64_t d, int32_t* id_fault_flag,
                       void* stream);
int tt_embedding_gather_f32(const float* table, const int64_t* ids, float* out, int64_t B,
                            int64_t d, int64_t vocab, void* stream);
int tt_embedding_gather_bf16(const float* table, const int64_t* ids, uint16_t* out, int64_t B,
                             int64_t d, int64_t vocab, void* stream);
int tt_embedding_bag_fwd(const float* table, const int64_t* values, const int64_t* offsets,
                         int32_t mode, void* out, int32_t out_dtype, int64_t num_bags, int64_t d,
                         int64_t vocab, void* stream);

/* ---------------------------------------------------------------------------------------
 * K5  sparse gradient -> dedup -> row-wise optimizer.
 * Replaces Keras optimizer._deduplicate_sparse_grad (tf.unique + unsorted_segment_sum) and
 * the sparse Adagrad/Adam update_step; learning_rate from
 * /root/reference/configs/data_config.yaml:63.
 * The gradient is given un-expanded: grad[num_rows, d] is d(loss)/d(tower input) and entry j
 * of `values` belongs to row bag(j) (offsets == NULL: bag(j) = j); mean pooling scales by
 * 1/len(bag).  first_flag (nullable, [nnz] bytes) is set to 1 where entry j is the first
 * occurrence of its id: values[first_flag] in order is exactly tf.unique(values).
 * The workspace must have been initialised once with tt_sparse_workspace_init; every
 * successful update leaves it clean.
 * ------------------------------------------------------------------------------------- */
int64_t tt_sparse_workspace_bytes(int64_t nnz, int64_t d);
int tt_sparse_workspace_init(void* workspace, int64_t workspace_bytes, int64_t nnz, int64_t d,
                             void* stream);
int tt_sparse_adagrad_update(float* table, float* accum, int64_t vocab, int64_t d,
                             const int64_t* values, const int64_t* offsets, int32_t mode,
                             int64_t num_rows, int64_t nnz, const float* grad, float lr, float eps,
                             void* workspace, int64_t workspace_bytes, uint8_t* first_flag,
                             void* stream);
/* LazyAdam (touched rows only; NOT Keras Adam).  alpha =
 * lr*sqrt(1-b2^t)/(1-b1^t) is computed by the caller. */
int tt_sparse_lazy_adam_update(float* table, float* m, float* v, int64_t vocab, int64_t d,
                               const int64_t* values, const int64_t* offsets, int32_t mode,
                               int64_t num_rows, int64_t nnz, const float* grad, float alpha,
                               float beta1, float beta2, float eps, void* workspace,
                               int64_t workspace_bytes, uint8_t* first_flag, void* stream);

/* Dense Keras Adagrad / Adam on a [rows, cols] variable.  grad_parts holds `num_parts`
 * stacked partial gradients [num_parts, rows, cols] (split-K partials of the wgrad GEMM) that
 * are summed in index order (deterministic).  l2 adds 2*l2*w (kernel_regularizer=l2,
 * /root/reference/configs/data_config.yaml:59).  shadow (nullable): bf16 copy [rows, cols] of
 * the updated variable, the operand the tensor-core kernels read. */
int tt_dense_adagrad_update(float* w, float* accum, const float* grad_parts, int32_t num_parts,
                            int64_t rows, int64_t cols, float lr, float eps, float l2,
                            uint16_t* shadow, void* stream);
int tt_dense_adam_update(float* w, float* m, float* v, const float* grad_parts, int32_t num_parts,
                         int64_t rows, int64_t cols, float alpha, float beta1, float beta2,
                         float eps, float l2, uint16_t* shadow, void* stream);

/* out[0] (+)= scale * sum(x^2) in a fixed order: the kernel_regularizer=l2 term of
 * tfrs.models.Model.train_step's regularization_loss (sum(model.losses)). */
int tt_sum_squares(const float* x, int64_t n, float scale, float* out, int32_t accumulate,
                   void* stream);

/* The same updates for several variables in ONE launch each (the whole optimizer step of a
 * two-tower model is launch-bound at the BASELINE sizes).  host_vars: HOST arrays. */
#define TT_MAX_DENSE_VARS 16
#define TT_MAX_SPARSE_VARS 8
typedef struct tt_dense_var {
  float* w;
  float* slot0;             /* Adagrad accumulator / Adam m */
  float* slot1;             /* Adam v (NULL for Adagrad) */
  const float* grad_parts;  /* [num_parts, n] */
  int64_t n;                /* rows * cols */
  int32_t num_parts;
  float l2;
  uint16_t* shadow;         /* nullable bf16 copy */
} tt_dense_var;
int tt_dense_adagrad_update_multi(const tt_dense_var* host_vars, int32_t num_vars, float lr, float eps,
                                  void* stream);
int tt_dense_adam_update_multi(const tt_dense_var* host_vars, int32_t num_vars, float alpha, float beta1,
                               float beta2, float eps, void* stream);
typedef struct tt_sparse_var {
  float* table;
  float* slot0;             /* Adagrad accumulator / LazyAdam m */
  float* slot1;             /* LazyAdam v (NULL for Adagrad) */
  int64_t vocab, d;
  const int64_t* values;
  const int64_t* offsets;   /* nullable */
  int64_t num_rows, nnz;
  const float* grad;        /* [num_rows, d] */
  void* workspace;
  int64_t workspace_bytes;
  uint8_t* first_flag;      /* nullable */
  int32_t mode;
  int32_t shard;            /* 0: `values` are rows of `table`.  (world << 16) | rank: `values` are GLOBAL ids of a
                             * table row-sharded over `world` ranks and `table` is this rank's shard: entries
                             * owned by other ranks (id % world != rank) are skipped, the others update local
                             * row id / world (tt_optimizer_prepare_sparse / tt_*_step only) */
} tt_sparse_var;
int tt_sparse_adagrad_update_multi(const tt_sparse_var* host_vars, int32_t num_vars, float lr, float eps,
                                   void* stream);
int tt_sparse_lazy_adam_update_multi(const tt_sparse_var* host_vars, int32_t num_vars, float alpha,
                                     float beta1, float beta2, float eps, void* stream);

/* The whole optimizer step in two launches.  tt_optimizer_prepare_sparse needs only the ids (table,
 * values, offsets, sizes, workspace of each tt_sparse_var): enqueue it when the lookup happens, on a
 * side stream, and it overlaps the forward pass.  tt_adagrad_step / tt_lazy_adam_step then update ALL
 * dense variables and ALL tables in one launch: an id that occurs once is applied straight from its
 * gradient row; duplicates are reduced into the first occurrence's accumulation row and the last arriver
 * applies (same results as tt_sparse_*_update: exact row set, fp32 sum order of duplicates not fixed).
 * Dense variables (Adam arithmetic for tt_lazy_adam_step) as tt_dense_*_update_multi.  The prepare call
 * must have completed (stream order / event) on the SAME workspaces before the step. */
int tt_optimizer_prepare_sparse(const tt_sparse_var* host_vars, int32_t num_vars, void* stream);
/* vars[i].w (fp32 [n]) = ordered sum of vars[i].grad_parts [num_parts, n], all variables in one launch:
 * the dense gradients are folded into one flat bucket before the data-parallel all-reduce. */
int tt_fold_parts_multi(const tt_dense_var* host_vars, int32_t num_vars, void* stream);
int tt_adagrad_step(const tt_dense_var* host_dense, int32_t num_dense, const tt_sparse_var* host_sparse,
                    int32_t num_sparse, float lr, float eps, void* stream);
/* alpha_device (nullable): device fp32 scalar that overrides `alpha` -- the bias-corrected step size of the current
 * iteration written by tt_adam_bias_correction, so that a captured CUDA graph applies a fresh alpha on every replay. */
int tt_lazy_adam_step(const tt_dense_var* host_dense, int32_t num_dense, const tt_sparse_var* host_sparse,
                      int32_t num_sparse, float alpha, const float* alpha_device, float beta1, float beta2, float eps,
                      void* stream);
/* Keras Adam step size: t = ++(*step_device); *alpha_device = lr * sqrt(1 - beta2^t) / (1 - beta1^t). */
int tt_adam_bias_correction(int64_t* step_device, float lr, float beta1, float beta2, float* alpha_device, void* stream);

/* ---------------------------------------------------------------------------------------
 * K2  tower MLP: tf.keras.layers.Dense forward / backward (layer sizes
 * /root/reference/configs/data_config.yaml:56-57).  kernel is ALWAYS the Keras layout
 * [in, out]; no transposed copy of any operand is needed (the tcgen05 descriptors read
 * K-major or MN-major shared-memory tiles as required).
 *
 * precision TT_F32 : x, kernel, y fp32 (CUDA-core FFMA path, 1e-5 parity).
 * precision TT_BF16: x bf16 [M,in], kernel bf16 [in,out] (the shadow copy), fp32 accumulation
 *                    on tcgen05; y bf16 [M,out]; y_f32 (nullable) fp32 copy.
 * ------------------------------------------------------------------------------------- */
int tt_dense_fwd(int32_t precision, const void* x, const void* kernel, const float* bias,
                 void* y, float* y_f32, int64_t M, int64_t in_dim, int64_t out_dim,
                 int32_t relu, void* stream);
/* dx = (dy @ kernel^T) [* (x > 0) when relu_mask_x != 0: x is the relu OUTPUT of the previous
 * layer, so the mask applies the previous layer's activation gradient];
 * dkernel_parts[p] = partial x^T @ dy over the p-th M-slice; dbias_parts[p] = colsum(dy) over
 * the same slice (nullable: the caller then takes tt_colsum_f32 of an fp32 copy of dy -- the
 * bias gradient is a cancelling sum, so it is taken before bf16 rounding whenever possible).
 * TT_F32 : everything fp32, num_parts must be 1.
 * TT_BF16: dy bf16 [M,out], x bf16 [M,in], kernel bf16 [in,out]; dx bf16 [M,in] (nullable);
 *          dx_f32 (nullable) fp32 copy of dx (the embedding gradient of the first layer);
 *          dkernel_parts fp32 [num_parts, in, out]; dbias_parts fp32 [num_parts, out]. */
int tt_dense_bwd(int32_t precision, const void* dy, const void* x, const void* kernel, void* dx,
                 float* dx_f32, float* dkernel_parts, int32_t num_parts, float* dbias_parts,
                 int64_t M, int64_t in_dim, int64_t out_dim, int32_t relu_mask_x, void* stream);
int32_t tt_dense_bwd_num_parts(int32_t precision, int64_t M, int64_t in_dim, int64_t out_dim);

/* out_parts[p, c] = sum over the p-th slice of rows of x[r, c], fixed order (fp32); the
 * num_parts partial rows feed tt_dense_*_update's grad_parts.  tt_sum_parts_f32 folds
 * stacked partials [num_parts, n] into one (before an all-reduce). */
int tt_colsum_f32(const float* x, float* out_parts, int64_t rows, int64_t cols, int32_t num_parts,
                  void* stream);
int tt_sum_parts_f32(const float* parts, int32_t num_parts, int64_t n, float* out, void* stream);

int tt_cast_f32_to_bf16(const float* in, uint16_t* out, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------
 * K1+K2 fused  tower forward / backward for the two-layer tower of the BASELINE configs
 * (Embedding [+ pooled multi-hot Embeddings] -> Dense(relu) -> Dense(linear); sizes
 * /root/reference/configs/data_config.yaml:55-57, ids /root/reference/src/data/preprocessor.py:481-489).
 * One launch covers up to TT_MAX_TOWERS towers (query + candidate): a CTA owns 128 batch rows of
 * one tower, gathers / pools them straight into the swizzled shared-memory A tile, and runs both
 * Dense layers on tcgen05 with the hidden activations kept on chip.  bf16 operands, fp32
 * accumulation (TT_BF16 precision of tt_dense_*).  Same arithmetic as tt_tower_input_fwd ->
 * tt_dense_fwd(relu) -> tt_dense_fwd, so the same oracle checks it.
 *
 * forward : x [B,d_in] = pooled tower input (bf16, saved for the backward; with num_feats == 0 x is
 *           an INPUT instead -- the rows a row-sharded lookup has already exchanged),
 *           h [B,d_hid] = relu(x w1 + b1) (bf16, saved), y [B,d_out] = h w2 + b2 (bf16).
 * backward: dy = sum of `dy_splits` stacked fp32 partials [dy_splits, B, d_out] (the split
 *           partials tt_retrieval_loss_bwd_parts leaves, or any fp32 gradient with splits = 1),
 *           dx [B,d_in] fp32 = the embedding-row gradient, and per 128-row slice p of the batch
 *           (num_parts = ceil(B / 128)): dw1_parts [P,d_in,d_hid], dw2_parts [P,d_hid,d_out],
 *           db1_parts [P,d_hid], db2_parts [P,d_out] (fp32; db2 from the un-rounded dy), to be
 *           summed in index order by tt_dense_*_update[_multi].
 * Supported shapes: tt_tower_mlp2_supported(d_in, d_hid, d_out) != 0 (d_in = 128, d_hid in
 * {128, 256}, d_out in {64, 128}); other shapes use the per-layer entry points.
 * ------------------------------------------------------------------------------------- */
#define TT_MAX_TOWERS 4
typedef struct tt_tower_mlp2 {
  tt_feature feats[TT_MAX_FEATURES];
  int32_t num_feats;
  int32_t d_in, d_hid, d_out;
  int64_t batch;
  const uint16_t* w1;   /* bf16 [d_in, d_hid]  (Keras layout) */
  const float* b1;      /* [d_hid] */
  const uint16_t* w2;   /* bf16 [d_hid, d_out] */
  const float* b2;      /* [d_out] */
  uint16_t* x;          /* bf16 [B, d_in] : forward output, backward input */
  uint16_t* h;          /* bf16 [B, d_hid]: forward output, backward input */
  uint16_t* y;          /* bf16 [B, d_out]: forward output */
  const float* dy_parts; /* backward: fp32 [dy_splits, B, d_out] */
  int32_t dy_splits;
  int32_t reserved;
  float* dx;            /* backward outputs */
  float* dw1_parts;
  float* dw2_parts;
  float* db1_parts;
  float* db2_parts;
  /* forward, optional: a tt_sparse_workspace_* buffer of the optimizer for this tower's table.  When non-NULL (the
   * tower must then have exactly one unsharded ID feature) the forward kernel also performs the hash insert of
   * tt_optimizer_prepare_sparse for the tower's ids -- hidden under its row gather -- so that the step needs neither
   * that launch nor a side stream for it; tt_adagrad_step / tt_lazy_adam_step then run on the same workspace. */
  void* prepare_workspace;
  int64_t prepare_workspace_bytes;
} tt_tower_mlp2;
int32_t tt_tower_mlp2_supported(int32_t d_in, int32_t d_hid, int32_t d_out);
int tt_tower_mlp2_fwd(const tt_tower_mlp2* host_towers, int32_t num_towers, int32_t* id_fault_flag,
                      void* stream);
int tt_tower_mlp2_bwd(const tt_tower_mlp2* host_towers, int32_t num_towers, void* stream);

/* ---------------------------------------------------------------------------------------
 * K3/K4  tfrs.tasks.Retrieval: in-batch softmax cross-entropy, logits never written to HBM.
 * Replaces matmul(q, c, transpose_b) [/ temperature] [- log clip(p)] [accidental-hit mask]
 * + CategoricalCrossentropy(from_logits=True, reduction=SUM) with labels = eye(nq, nc)
 * (in_batch + temperature per /root/reference/configs/data_config.yaml:68-70).
 *
 * q [nq, d], c [nc, d] (fp32 for TT_F32, bf16 for TT_BF16).  Row i's positive is column
 * label_offset + i (label_offset = rank * nq under all-gathered candidates).
 * Optional: sample_weight [nq] fp32; cand_log_q [nc] fp32 = log(clip(p, 1e-6, 1)) subtracted
 * from every row; accidental-hit removal when cand_ids != NULL (int64 [nc]): column j of row
 * i gets MIN_FLOAT added when cand_ids[j] == cand_ids[label(i)] and j != label(i).
 * Outputs: row_lse [nq], row_pos [nq] (the transformed positive logit), loss[1] =
 * sum_i w_i (lse_i - pos_i) reduced in a fixed order.
 * ------------------------------------------------------------------------------------- */
int64_t tt_retrieval_workspace_bytes(int32_t precision, int64_t nq, int64_t nc, int64_t d);
/* TT_BF16: the head of the workspace holds arrival tickets (the forward's last CTA per row block
 * folds the candidate splits and the loss in-kernel).  Call this ONCE after allocating a workspace
 * (it zeroes the tickets; equivalently zero-fill the buffer); every launch leaves them zero again.
 * A workspace must not be shared by launches that may run concurrently. */
int tt_retrieval_workspace_init(int32_t precision, void* workspace, int64_t workspace_bytes, int64_t nq,
                                int64_t nc, int64_t d, void* stream);
int tt_retrieval_loss_fwd(int32_t precision, const void* q, const void* c, int64_t nq, int64_t nc,
                          int64_t d, float inv_temperature, int64_t label_offset,
                          const float* sample_weight, const float* cand_log_q,
                          const int64_t* cand_ids, float* row_lse, float* row_pos, float* loss,
                          void* workspace, int64_t workspace_bytes, void* stream);
/* dq [nq, d] fp32, dc [nc, d] fp32 (this rank's partial when candidates are all-gathered).
 * TT_BF16 can also emit bf16 copies dq_bf16 [nq,d] / dc_bf16 [nc,d] (nullable) for the MLP
 * backward that follows.  grad_scale multiplies the upstream gradient (1.0 for SUM). */
int tt_retrieval_loss_bwd(int32_t precision, const void* q, const void* c, int64_t nq, int64_t nc,
                          int64_t d, float inv_temperature, int64_t label_offset,
                          const float* sample_weight, const float* cand_log_q,
                          const int64_t* cand_ids, const float* row_lse, float grad_scale,
                          float* dq, float* dc, uint16_t* dq_bf16, uint16_t* dc_bf16,
                          void* workspace, int64_t workspace_bytes, void* stream);

/* Split form of the backward for a consumer that folds the partial sums itself
 * (tt_tower_mlp2_bwd): the dQ pass leaves *dq_splits stacked fp32 partials [dq_splits, nq, d] in
 * dq_parts, the dC pass [dc_splits, nc, d] in dc_parts; dq = sum over splits in index order.
 * Buffers are caller-owned, sized with tt_retrieval_bwd_num_splits.  TT_BF16 only.
 * tt_combine_parts_f32 is that ordered sum as a stand-alone kernel (fp32 and/or bf16 output). */
/* tfrs.tasks.Retrieval(num_hard_negatives = n) (tfrs layers/loss.py HardNegativeMining): the loss over the positive and
 * the n highest-scoring negatives of every query row.  `selected` int64 [nq, k_sel]: column 0 = index of the positive
 * candidate, then the selected negatives (-1 = unused slot) -- built by the caller from tt_topk_bruteforce with
 * k = n + 1 (tf.math.top_k order).  Evaluates the gathered logits only: scores [nq, k_sel] fp32 (scaled by
 * inv_temperature; -inf for unused slots), row_lse, row_pos, row_loss [nq] and the SUM loss [1] (fixed summation order).
 * Backward: dq [nq, d] and dc [nc, d] fp32 (dc is zeroed, then accumulated with fp32 atomics).  q / c: fp32 (TT_F32) or
 * bf16 (TT_BF16), d <= 256, arithmetic in fp32. */
int tt_hard_negative_loss_fwd(int32_t precision, const void* q, const void* c, const int64_t* selected, int64_t nq,
                              int64_t nc, int64_t k_sel, int64_t d, float inv_temperature, const float* sample_weight,
                              float* scores, float* row_lse, float* row_pos, float* row_loss, float* loss, void* stream);
int tt_hard_negative_loss_bwd(int32_t precision, const void* q, const void* c, const int64_t* selected, int64_t nq,
                              int64_t nc, int64_t k_sel, int64_t d, float inv_temperature, float grad_scale,
                              const float* sample_weight, const float* scores, const float* row_lse, float* dq,
                              float* dc, void* stream);

/* Forward + dQ in one pass (TT_BF16, d = 64 or 128, no log-q correction / accidental-hit mask): the loss forward
 * with a second GEMM per tile that accumulates sum_j exp(s_ij - m_i) c_j (flash-attention forward shape), so
 * dq = (w_i / T) (softmax(S) C - c_label) -- d(loss)/dq for an upstream gradient of 1 -- comes out of the same pass
 * and the backward needs the dC pass only (tt_retrieval_loss_bwd_parts with dq_parts = NULL).  Outputs as
 * tt_retrieval_loss_fwd plus dq fp32 [nq, d].  The workspace (tt_retrieval_fwd_dq_workspace_bytes; 0 = shape not
 * supported) must be zero in its first 256 bytes before the first call; the call leaves it so.
 * finalize_stream (nullable): the final summation of the scalar loss (one block; nothing on the device waits for it)
 * is forked onto that stream (event record + wait inside the call; the CALLER joins it back before reading `loss`);
 * row_lse and dq are produced on `stream`.
 * tt_retrieval_loss_bwd_dc_fused: the dC pass for this forward when there are no sample weights (must be NULL).  It
 * reads -lse per query column (log2 domain, written into the workspace by the fold) with broadcast loads instead of
 * staging row_lse through shared memory; dc_parts as in tt_retrieval_loss_bwd_parts. */
int64_t tt_retrieval_fwd_dq_workspace_bytes(int64_t nq, int64_t nc, int64_t d);
int tt_retrieval_loss_fwd_dq(const void* q, const void* c, int64_t nq, int64_t nc, int64_t d,
                             float inv_temperature, int64_t label_offset, const float* sample_weight,
                             float* row_lse, float* row_pos, float* loss, float* dq, void* workspace,
                             int64_t workspace_bytes, void* stream, void* finalize_stream);
int tt_retrieval_loss_bwd_dc_fused(const void* q, const void* c, int64_t nq, int64_t nc, int64_t d,
                                   float inv_temperature, int64_t label_offset, const float* sample_weight,
                                   const void* fwd_dq_workspace, float grad_scale, float* dc_parts, void* stream);
int tt_retrieval_bwd_num_splits(int32_t precision, int64_t nq, int64_t nc, int64_t d, int32_t* dq_splits,
                                int32_t* dc_splits);
int tt_retrieval_loss_bwd_parts(int32_t precision, const void* q, const void* c, int64_t nq, int64_t nc,
                                int64_t d, float inv_temperature, int64_t label_offset,
                                const float* sample_weight, const float* cand_log_q,
                                const int64_t* cand_ids, const float* row_lse, float grad_scale,
                                float* dq_parts, float* dc_parts, void* stream);
int tt_combine_parts_f32(const float* parts, int32_t num_parts, int64_t rows, int64_t d, float* out_f32,
                         uint16_t* out_bf16, void* stream);

/* ---------------------------------------------------------------------------------------
 * K6  brute-force scoring + top-k.
 * Replaces tfrs.layers.factorized_top_k.BruteForce.call (matmul + tf.math.top_k) and
 * faiss.IndexFlatIP.search (ks from
 * /root/reference/configs/data_config.yaml:71).  Result rows are sorted by score descending,
 * ties by LOWER candidate index first (the tf.math.top_k rule).
 * out_ids[q, j] = identifiers ? identifiers[idx] : cand_index_base + idx.
 * The candidate range may be split over `num_splits` CTAs columns (workspace holds the
 * partial lists) and merged with the same ordering rule.
 * Two stages, so that the ids do not depend on the accumulation order of the scoring kernel:
 * the scoring stage keeps the best k + TT_TOPK_RERANK_MARGIN candidates per query, the second
 * stage recomputes their scores exactly (fp64 sum of the exact products), rounds once to fp32
 * (the dtype of the TFRS score tensor) and orders by (score desc, index asc).  out_scores are
 * those correctly rounded scores.  uncertain_rows (device int32, optional, incremented): number
 * of query rows for which the margin could not be shown to be wide enough (expected 0).
 * workspace_bytes >= tt_topk_workspace_bytes(...) is always required.
 * bf16 scoring stage, d <= 128 and enough candidates for a sample (csrc/topk_scan.cu): a strided ~1/32 sample of the
 * candidate tiles gives every query row a score threshold that at least k + margin candidates are guaranteed to
 * reach; the scan then only compares against it and the few thousand survivors per row are sorted by one CTA.  A row
 * whose survivor buffer overflows raises a device flag and the list-keeping kernel redoes the batch (same result).
 * tt_topk_num_launches: kernels one tt_topk_bruteforce call enqueues for this shape.
 * ------------------------------------------------------------------------------------- */
int32_t tt_topk_num_splits(int32_t precision, int64_t nq, int64_t nc, int64_t d, int32_t k);
int32_t tt_topk_num_launches(int32_t precision, int64_t nq, int64_t nc, int64_t d, int32_t k);
int64_t tt_topk_workspace_bytes(int32_t precision, int64_t nq, int64_t nc, int64_t d, int32_t k);
int tt_topk_bruteforce(int32_t precision, const void* queries, const void* candidates, int64_t nq,
                       int64_t nc, int64_t d, int32_t k, int64_t cand_index_base,
                       const int64_t* identifiers, float* out_scores, int64_t* out_ids,
                       int32_t* uncertain_rows, void* workspace, int64_t workspace_bytes, void* stream);
/* Candidate-sharded serving (BASELINE configs[4]): every rank scores ALL nq queries against ITS shard
 * (global index of local candidate 0 = cand_index_base) and the exact partial top-k of query qi is written by the kernel's
 * epilogue -- posted NVLink stores, no exchange kernel -- into the symmetric workspace of the rank that merges that query,
 * owner = qi / queries_per_rank: scores [world, queries_per_rank, k] fp32 at peer_bases[owner] + recv_scores_offset,
 * global candidate indices [world, queries_per_rank, k] int64 at + recv_ids_offset, list slot [rank].  After a
 * tt_peer_barrier the owner runs tt_topk_merge over its `world` lists.  peer_bases: device int64 [world]. */
int tt_topk_bruteforce_peer(int32_t precision, const void* queries, const void* candidates, int64_t nq,
                            int64_t nc, int64_t d, int32_t k, int64_t cand_index_base,
                            const int64_t* peer_bases, int32_t world, int32_t rank, int64_t queries_per_rank,
                            int64_t recv_scores_offset, int64_t recv_ids_offset, int32_t* uncertain_rows,
                            void* workspace, int64_t workspace_bytes, void* stream);
/* Merge `num_lists` sorted top-k lists per query: scores/ids are [num_lists, nq, k_in]
 * (list-major, as written by per-shard searches after an all-gather); ids are candidate
 * INDICES and the ordering rule is (score desc, index asc).  Writes [nq, k_out] with
 * out_ids = identifiers ? identifiers[index] : index_base + index. */
int tt_topk_merge(const float* scores, const int64_t* ids, int32_t num_lists, int64_t nq,
                  int32_t k_in, int32_t k_out, int64_t index_base, const int64_t* identifiers,
                  float* out_scores, int64_t* out_ids, void* stream);
/* FactorizedTopK hit counting.  Score mode (true_ids == NULL): hit@k iff
 * fewer than k of topk_scores[i,:] are strictly greater than positive[i].  Id mode: hit@k iff
 * true_ids[i] appears in topk_ids[i,:k].  hits_out[j] += sum_i w_i * hit_{ks[j]}(i);
 * weight_out[0] += sum_i w_i.  host_ks: HOST array of num_ks ints (<= 8). */
int tt_topk_hits(const float* positive, const float* topk_scores, const int64_t* topk_ids,
                 const int64_t* true_ids, const float* sample_weight, int64_t nq, int32_t k,
                 const int32_t* host_ks, int32_t num_ks, float* hits_out, float* weight_out,
                 void* stream);
/* positive[i] = <q[i,:], c[i,:]> in fp32 (inputs of `precision`). */
int tt_rowwise_dot(int32_t precision, const void* q, const void* c, float* out, int64_t n,
                   int64_t d, void* stream);

/* ---------------------------------------------------------------------------------------
 * Row-sharded tables across ranks: stable partition of ids by owner.
 * owner(id) = id % world (cyclic).  Outputs: perm [n] (position of entry j in the
 * owner-major send buffer), send_ids = local row (id / world) in owner-major order,
 * counts [world].  Stable: entries of one owner keep their batch order, so per-owner
 * buckets are bit-reproducible.  capacity == 0: buckets are packed (send_ids [n]);
 * capacity > 0: bucket o starts at o * capacity (send_ids [world * capacity], unused slots
 * -1) so the all-to-all has static shapes; an entry that does not fit gets perm = -1 and
 * *overflow_flag = 1 (nullable).
 * ------------------------------------------------------------------------------------- */
int tt_partition_ids(const int64_t* ids, int64_t n, int32_t world, int64_t capacity,
                     int64_t* send_ids, int64_t* perm, int64_t* counts, int32_t* overflow_flag,
                     void* stream);
/* Rows of row_bytes (multiple of 16): out[perm[j]] = in[j] (inverse = 0) or out[j] = in[perm[j]]
 * (inverse = 1); entries with perm[j] < 0 are skipped. */
int tt_permute_rows(const void* in, const int64_t* perm, void* out, int64_t n, int64_t row_bytes,
                    int32_t inverse, void* stream);

/* ---------------------------------------------------------------------------------------
 * NVLink peer-memory exchange (csrc/peer.cu): the collectives of the row-sharded step as kernels on a
 * SYMMETRIC workspace -- same layout on every rank, every rank maps all copies.  `bases` / `flag_bases`
 * are DEVICE arrays [world] of the peer-mapped addresses of the workspace / of its flag block (8 slots x 16
 * uint64, zeroed once); `step_counter` is a device uint64[16] of this rank: [0, 8) the epoch of each slot
 * (initialised to 1), [8, 16) zero.  The k-th barrier on a slot signals k into that slot of every rank, waits until
 * every rank has signalled >= k, and advances the slot's epoch.  They replace, one for one:
 *   tt_peer_push      all-gather (producer writes its block into every rank's copy; up to 4 segments)
 *   tt_peer_combine_scatter  reduce-scatter, producer side: row i of the ordered sum of `splits` stacked partials
 *                     [splits, world * rows_per_rank, d] goes to slot [rank] of the [world, rows_per_rank, d]
 *                     receive area (dst_offset) of rank i / rows_per_rank; the owner folds the slots
 *   tt_peer_push_rows gradient row j of table t -> row [rank * b + j] of the [world * b, d] area (dst_offset[t]) of
 *                     the rank that owns table row ids[t][j] (id % world); rows of other owners stay untouched there
 *   tt_peer_sum_f32   all-reduce: out[i] = sum_r source_r[offset + r * stride + i], added in rank order on every
 *                     rank (bit-identical replicas); slot >= 0 runs a barrier first, slot < 0 none
 *   tt_peer_pull_rows owner-side gather of the gradient rows of a row-sharded table: for every global batch
 *                     position j (produced by rank j / rows_per_rank) whose id this rank owns (id % world ==
 *                     rank), out[t][j, :] = that rank's rows at src_offset[t]; other rows of out are untouched
 * No NCCL inside; torch.distributed only allocates / rendezvouses the workspace.
 * ------------------------------------------------------------------------------------- */
int tt_peer_barrier(const void* flag_bases, void* step_counter, int32_t world, int32_t rank, int32_t slot,
                    void* stream);
int tt_peer_push(const void* bases, int32_t world, int32_t nseg, const void* const* src,
                 const int64_t* dst_offset, const int64_t* bytes, void* stream);
int tt_peer_sum_f32(const void* bases, int64_t offset_bytes, int64_t stride_bytes, int64_t n, float* out,
                    const void* flag_bases, void* step_counter, int32_t world, int32_t rank, int32_t slot, void* stream);
int tt_peer_combine_scatter(const void* bases, const float* parts, int32_t splits, int64_t rows, int64_t d,
                            int64_t rows_per_rank, int64_t dst_offset, int32_t world, int32_t rank, void* stream);
/* The reduce-scatter of dC fused into the dC pass (no combine kernel): `peer_maps` is a DEVICE array [world] of
 * 128-byte tensor maps, 64-byte aligned, built on the host by tt_peer_make_row_maps over every rank's
 * [world * b, d] fp32 receive area (peer-mapped addresses) and copied to the device once.  The kernel's epilogue
 * TMA-stores each 128-row block of dC into slot [rank] of the rank that owns those candidates while other CTAs still
 * compute.  Needs b = nc / world a multiple of 128, d <= 128, no log-q / accidental-hit terms, and an unsplit dC pass
 * (more 128-row blocks of candidates than half the SMs; else error -> use tt_retrieval_loss_bwd_parts + tt_peer_combine_scatter).  `scratch`: fp32
 * [nc, d] caller-owned (only used for ragged blocks, i.e. never when the conditions hold). */
int tt_peer_make_row_maps(const uint64_t* peer_addrs, int32_t world, int64_t rows, int64_t d, void* out_host);
int tt_peer_retrieval_bwd_dc(const void* q, const void* c, int64_t nq, int64_t nc, int64_t d, float inv_temperature,
                             int64_t label_offset, const float* sample_weight, const float* row_lse,
                             float grad_scale, const void* peer_maps, int32_t world, int32_t rank,
                             float* scratch, void* stream);
int tt_peer_push_rows(const void* bases, int32_t ntab, const int64_t* const* ids, const float* const* src,
                      const int64_t* dst_offset, int64_t b, int64_t d, int32_t world, int32_t rank, void* stream);
int tt_peer_pull_rows(const void* bases, int32_t ntab, const int64_t* const* ids, const int64_t* src_offset,
                      float* const* out, int64_t rows_per_rank, int64_t d, const void* flag_bases,
                      void* step_counter, int32_t world, int32_t rank, int32_t slot, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TWOTOWER_H_ */
