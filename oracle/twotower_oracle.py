"""numpy restatement of the TFRS / Keras / tf.math.top_k / FAISS-flat arithmetic on the
two-tower hot path.  TEST INFRASTRUCTURE ONLY -- never imported by the product package.

PARITY UNPINNED.  The reference tree holds no implementation of this path
(/root/reference/src/models/__init__.py:1, src/training/__init__.py:1,
src/serving/__init__.py:1, src/evaluation/__init__.py:1 are one-line docstrings) and its
tests pin nothing at this boundary (/root/reference/tests/unit/* cover src/data only).
The arithmetic lives in un-vendored dependencies with lower-bound pins only
(/root/reference/pyproject.toml:22 tensorflow>=2.15.0, :24 tensorflow-recommenders>=0.7.3,
:39 faiss-cpu>=1.7.4).  Each function below names the upstream routine it restates
(SURVEY.md Appendix A.x) and the in-repo line that parameterises it.

Every routine takes a ``dtype`` (np.float64 = truth, np.float32 = reference precision).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "MIN_FLOAT", "MAX_FLOAT", "bf16_round", "embedding_lookup", "embedding_bag",
    "embedding_bag_backward", "dense_forward", "mlp_forward", "mlp_backward",
    "l2_regularization", "retrieval_scores", "retrieval_loss", "retrieval_loss_and_grads",
    "dedup_sparse_grad", "adagrad_dense", "adagrad_sparse", "adam_dense", "adam_sparse_keras",
    "lazy_adam_sparse", "top_k", "brute_force_topk", "streaming_topk", "topk_merge",
    "factorized_topk_hits", "FactorizedTopKOracle", "keras_uniform", "glorot_uniform",
    "TowerSpec", "tower_forward", "tower_backward", "two_tower_train_step", "init_tower",
]

# tfrs/tasks/retrieval.py + tfrs/layers/loss.py: MIN_FLOAT = np.finfo(np.float32).min / 100.0
MIN_FLOAT = float(np.finfo(np.float32).min / 100.0)
MAX_FLOAT = float(np.finfo(np.float32).max / 100.0)


# ----------------------------------------------------------------------------------------
# precision helpers
# ----------------------------------------------------------------------------------------
def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round fp32 -> bf16 (round-to-nearest-even) and return as fp32.  Used to emulate the
    bf16 storage points of the tensor-core path (inputs of every MMA are bf16, accumulation
    is fp32)."""
    a = np.ascontiguousarray(x, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    rounding = ((u >> 16) & 1) + 0x7FFF
    r = ((u + rounding) >> 16) << 16
    out = r.astype(np.uint32).view(np.float32).reshape(a.shape)
    nan = np.isnan(a)
    if nan.any():
        out = out.copy()
        out[nan] = np.nan
    return out


# ----------------------------------------------------------------------------------------
# A.3 towers: Embedding / pooled multi-hot / Dense
# ----------------------------------------------------------------------------------------
def keras_uniform(rng: np.random.Generator, shape, dtype=np.float32) -> np.ndarray:
    """Keras ``Embedding`` default initializer "uniform" = RandomUniform(-0.05, 0.05)."""
    return rng.uniform(-0.05, 0.05, size=shape).astype(dtype)


def glorot_uniform(rng: np.random.Generator, fan_in: int, fan_out: int, dtype=np.float32):
    """Keras ``Dense`` default kernel initializer; kernel shape [fan_in, fan_out]."""
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(dtype)


def embedding_lookup(table: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """tf.keras.layers.Embedding forward = tf.nn.embedding_lookup: a pure gather.
    Ids are the dense int64 codes the data layer emits
    (/root/reference/src/data/preprocessor.py:481-482).  Out-of-range ids are an error on
    CPU in TF; same here."""
    ids = np.asarray(ids)
    if ids.size and (ids.min() < 0 or ids.max() >= table.shape[0]):
        raise IndexError("embedding id out of range")
    return table[ids]


def embedding_bag(table: np.ndarray, values: np.ndarray, offsets: np.ndarray, mode: str = "mean",
                  dtype=np.float64) -> np.ndarray:
    """Multi-hot pooling = safe_embedding_lookup_sparse(combiner=mode) over CSR bags.
    mean = sum / bag length, duplicates inside a bag count each time, empty bag -> zeros."""
    nb = len(offsets) - 1
    d = table.shape[1]
    out = np.zeros((nb, d), dtype=dtype)
    for b in range(nb):
        lo, hi = int(offsets[b]), int(offsets[b + 1])
        if hi > lo:
            acc = np.zeros(d, dtype=dtype)
            for j in range(lo, hi):          # sequential order = the kernel's order
                acc = acc + table[values[j]].astype(dtype)
            out[b] = acc / (hi - lo) if mode == "mean" else acc
    return out


def embedding_bag_backward(values, offsets, dout: np.ndarray, mode: str = "mean"):
    """Gradient of embedding_bag wrt the table as IndexedSlices (indices with duplicates,
    one value row per member): each member row receives upstream / L (mean) or upstream."""
    nnz = int(offsets[-1])
    rows = np.zeros((nnz, dout.shape[1]), dtype=dout.dtype)
    for b in range(len(offsets) - 1):
        lo, hi = int(offsets[b]), int(offsets[b + 1])
        if hi > lo:
            rows[lo:hi] = dout[b] / (hi - lo) if mode == "mean" else dout[b]
    return np.asarray(values[:nnz]), rows


def dense_forward(x, kernel, bias, activation=None):
    """tf.keras.layers.Dense: activation(x @ kernel + bias), kernel [in, out]."""
    y = x @ kernel + bias
    if activation == "relu":
        y = np.maximum(y, 0)
    elif activation not in (None, "linear"):
        raise ValueError(activation)
    return y


def mlp_forward(x, kernels, biases, bf16: bool = False):
    """TFRS deep-tower pattern: Dense(u, relu) for all but the last layer, last layer linear
    (layer sizes from /root/reference/configs/data_config.yaml:56-57).
    Returns (out, activations) where activations[l] is the INPUT of layer l.
    bf16=True emulates the tensor-core path: inputs of each matmul rounded to bf16."""
    acts = []
    h = x
    n = len(kernels)
    for l, (w, b) in enumerate(zip(kernels, biases)):
        if bf16:
            h = bf16_round(h).astype(x.dtype)
            w = bf16_round(w).astype(x.dtype)
        acts.append(h)
        h = dense_forward(h, w, b, "relu" if l < n - 1 else None)
    return h, acts


def mlp_backward(dout, kernels, acts, out_last=None):
    """Autodiff of mlp_forward.  Returns (dx, dkernels, dbiases)."""
    n = len(kernels)
    dk = [None] * n
    db = [None] * n
    g = dout
    for l in range(n - 1, -1, -1):
        x = acts[l]
        if l < n - 1:
            # relu mask from the layer's OUTPUT = input of layer l+1
            g = g * (acts[l + 1] > 0)
        dk[l] = x.T @ g
        db[l] = g.sum(axis=0)
        g = g @ kernels[l].T
    return g, dk, db


def l2_regularization(kernels, lam):
    """kernel_regularizer=l2(lam): lam * sum(w^2) per Dense kernel, added to model.losses
    (/root/reference/configs/data_config.yaml:59)."""
    return float(sum(lam * np.sum(np.square(w.astype(np.float64))) for w in kernels))


# ----------------------------------------------------------------------------------------
# A.2 tfrs.tasks.Retrieval
# ----------------------------------------------------------------------------------------
def retrieval_scores(q, c, temperature=None, candidate_sampling_probability=None,
                     candidate_ids=None, remove_accidental_hits=False, dtype=np.float64):
    """scores after every transform of tfrs.tasks.Retrieval.call *before* hard-negative
    mining, in upstream order: matmul -> /temperature -> -log(clip(p,1e-6,1)) ->
    accidental-hit mask (+ MIN_FLOAT on non-diagonal duplicates of the row's positive id)."""
    q = q.astype(dtype)
    c = c.astype(dtype)
    nq, nc = q.shape[0], c.shape[0]
    s = q @ c.T
    if temperature is not None:
        s = s / dtype(temperature)
    if candidate_sampling_probability is not None:
        p = np.clip(np.asarray(candidate_sampling_probability, dtype=dtype), 1e-6, 1.0)
        s = s - np.log(p)[None, :]
    if remove_accidental_hits:
        if candidate_ids is None:
            raise ValueError("When accidental hit removal is enabled, candidate ids must be supplied.")
        ids = np.asarray(candidate_ids)
        pos = ids[:nq]                              # argmax(eye) = i
        dup = (pos[:, None] == ids[None, :]).astype(dtype)
        dup[np.arange(nq), np.arange(nq)] -= 1.0   # minus labels
        s = s + dup * dtype(MIN_FLOAT)
    return s


def _row_lse(s):
    m = s.max(axis=1, keepdims=True)
    return (m + np.log(np.exp(s - m).sum(axis=1, keepdims=True)))[:, 0]


def retrieval_loss_and_grads(q, c, temperature=None, sample_weight=None,
                             candidate_sampling_probability=None, candidate_ids=None,
                             remove_accidental_hits=False, num_hard_negatives=None,
                             dtype=np.float64):
    """tfrs.tasks.Retrieval.call with the default loss
    CategoricalCrossentropy(from_logits=True, reduction=SUM), labels = eye(nq, nc)
    (in_batch sampling + temperature per /root/reference/configs/data_config.yaml:68-70).
    Returns dict(loss, dq, dc, lse, pos, scores)."""
    nq, nc = q.shape[0], c.shape[0]
    s = retrieval_scores(q, c, temperature, candidate_sampling_probability, candidate_ids,
                         remove_accidental_hits, dtype)
    keep = np.ones_like(s, dtype=bool)
    if num_hard_negatives is not None:
        k = min(int(num_hard_negatives) + 1, nc)
        boosted = s.copy()
        boosted[np.arange(nq), np.arange(nq)] += dtype(MAX_FLOAT)
        keep[:] = False
        # top-k per row, ties -> lower index (tf.math.top_k rule)
        order = np.argsort(-boosted, axis=1, kind="stable")[:, :k]
        np.put_along_axis(keep, order, True, axis=1)
    sm = np.where(keep, s, -np.inf)
    lse = _row_lse(sm)
    pos = s[np.arange(nq), np.arange(nq)]
    w = np.ones(nq, dtype=dtype) if sample_weight is None else np.asarray(sample_weight, dtype=dtype)
    loss = float(np.sum(w * (lse - pos)))
    p = np.exp(sm - lse[:, None])
    ds = p
    ds[np.arange(nq), np.arange(nq)] -= 1.0
    ds = ds * w[:, None]
    if temperature is not None:
        ds = ds / dtype(temperature)
    dq = ds @ c.astype(dtype)
    dc = ds.T @ q.astype(dtype)
    return dict(loss=loss, dq=dq, dc=dc, lse=lse, pos=pos, scores=s)


def retrieval_loss(q, c, **kw):
    return retrieval_loss_and_grads(q, c, **kw)["loss"]


# ----------------------------------------------------------------------------------------
# A.6 Keras 2.15 optimizers on IndexedSlices
# ----------------------------------------------------------------------------------------
def dedup_sparse_grad(ids: np.ndarray, rows: np.ndarray):
    """Keras optimizer._deduplicate_sparse_grad: tf.unique (FIRST-OCCURRENCE order) +
    unsorted_segment_sum.  Returns (unique_ids, summed_rows, first_positions)."""
    ids = np.asarray(ids)
    uniq_sorted, first_idx, inv = np.unique(ids, return_index=True, return_inverse=True)
    order = np.argsort(first_idx, kind="stable")           # first-occurrence order
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    seg = rank[inv]
    summed = np.zeros((len(order), rows.shape[1]), dtype=rows.dtype)
    np.add.at(summed, seg, rows)
    return uniq_sorted[order], summed, first_idx[order]


def adagrad_dense(w, acc, g, lr=0.001, eps=1e-7):
    """Keras Adagrad.update_step (dense): acc += g^2; w -= lr * g / sqrt(acc + eps).
    learning_rate default from /root/reference/configs/data_config.yaml:63."""
    acc = acc + g * g
    w = w - lr * g / np.sqrt(acc + eps)
    return w, acc


def adagrad_sparse(table, acc, ids, rows, lr=0.001, eps=1e-7, inplace=False):
    """Keras Adagrad on IndexedSlices: dedup, then the row-wise update on touched rows only.
    Mutates copies unless inplace (the timing baseline updates in place, as TF's
    ResourceScatter ops do); returns (table, acc, unique_ids)."""
    if not inplace:
        table = table.copy()
        acc = acc.copy()
    u, g, _ = dedup_sparse_grad(ids, rows)
    g = g.astype(table.dtype)
    acc[u] = acc[u] + g * g
    table[u] = table[u] - table.dtype.type(lr) * g / np.sqrt(acc[u] + table.dtype.type(eps))
    return table, acc, u


def _one_minus_f32(b):
    """Keras casts beta to the variable dtype (fp32) before forming 1 - beta, so the factor
    carries fp32 rounding (1 - 0.999f = 0.00100005): restated exactly."""
    return float(np.float32(1.0) - np.float32(b))


def adam_dense(w, m, v, g, step, lr=0.001, b1=0.9, b2=0.999, eps=1e-7):
    """Keras 2.15 Adam.update_step (dense).  ``step`` is 1-based (iterations + 1)."""
    alpha = lr * np.sqrt(1.0 - b2 ** step) / (1.0 - b1 ** step)
    m = m + (g - m) * _one_minus_f32(b1)
    v = v + (g * g - v) * _one_minus_f32(b2)
    w = w - m * alpha / (np.sqrt(v) + eps)
    return w, m, v


def adam_sparse_keras(table, m, v, ids, rows, step, lr=0.001, b1=0.9, b2=0.999, eps=1e-7):
    """Keras 2.15 Adam on IndexedSlices: decays m, v of the WHOLE variable, scatter-adds
    the touched rows, updates ALL rows (== dense Adam with a scattered gradient)."""
    u, g, _ = dedup_sparse_grad(ids, rows)
    dense = np.zeros_like(table)
    dense[u] = g.astype(table.dtype)
    return adam_dense(table, m, v, dense, step, lr, b1, b2, eps)


def lazy_adam_sparse(table, m, v, ids, rows, step, lr=0.001, b1=0.9, b2=0.999, eps=1e-7):
    """TF-Addons LazyAdam: same arithmetic but on touched rows only (explicitly NOT Keras
    Adam; offered as the row-wise variant north_star item (4) names)."""
    table, m, v = table.copy(), m.copy(), v.copy()
    u, g, _ = dedup_sparse_grad(ids, rows)
    g = g.astype(table.dtype)
    alpha = lr * np.sqrt(1.0 - b2 ** step) / (1.0 - b1 ** step)
    m[u] = m[u] + (g - m[u]) * _one_minus_f32(b1)
    v[u] = v[u] + (g * g - v[u]) * _one_minus_f32(b2)
    table[u] = table[u] - m[u] * alpha / (np.sqrt(v[u]) + eps)
    return table, m, v, u


# ----------------------------------------------------------------------------------------
# A.4 BruteForce / Streaming / tf.math.top_k ; A.7 FAISS flat
# ----------------------------------------------------------------------------------------
def top_k(scores: np.ndarray, k: int):
    """tf.math.top_k: sorted descending; among equal values the LOWER index comes first."""
    order = np.argsort(-scores, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(scores, order, axis=1), order


def brute_force_topk(queries, candidates, k, identifiers=None, dtype=np.float64, block=4096, score_dtype=None):
    """tfrs.layers.factorized_top_k.BruteForce.call == faiss.IndexFlatIP.search on tie-free
    data: scores = q @ cand^T, top_k, ids = gather(identifiers, indices).  Blocked over
    candidates (what IndexFlatIP does with sgemm blocks + heaps), merging with the
    (score desc, index asc) rule so the result equals the unblocked tf.math.top_k.

    score_dtype: dtype of the score TENSOR that top_k orders.  TFRS orders an fp32 matmul result whose last
    bits depend on the BLAS summation order; the implementation-independent statement of that is
    dtype=float64, score_dtype=float32: the exact dot product (fp64 sum of exact products) rounded ONCE to
    fp32, ordered by (score desc, index asc).  None: order the `dtype` scores themselves."""
    q = queries.astype(dtype)
    nq = q.shape[0]
    n = candidates.shape[0]
    k = min(k, n)
    sd = dtype if score_dtype is None else score_dtype
    best_s = np.full((nq, 0), 0, dtype=sd)
    best_i = np.zeros((nq, 0), dtype=np.int64)
    for lo in range(0, n, block):
        hi = min(n, lo + block)
        s = (q @ candidates[lo:hi].astype(dtype).T).astype(sd)
        if hi - lo > 4 * k and best_s.shape[1] == k:
            best_s, best_i = _merge_block_filtered(best_s, best_i, s, lo, k)
            continue
        cs = np.concatenate([best_s, s], axis=1)
        ci = np.concatenate([best_i, np.broadcast_to(np.arange(lo, hi), (nq, hi - lo))], axis=1)
        # existing entries have lower indices and come first -> stable sort keeps the rule
        order = np.argsort(-cs, axis=1, kind="stable")[:, :k]
        best_s = np.take_along_axis(cs, order, axis=1)
        best_i = np.take_along_axis(ci, order, axis=1)
    ids = best_i if identifiers is None else np.asarray(identifiers)[best_i]
    return best_s, ids


def _merge_block_filtered(best_s, best_i, s, lo, k):
    """Same result as the concat + stable argsort above, for wide blocks: only block entries that are >= the
    row's current k-th best can enter (an equal score with a higher index never displaces a kept entry, but it
    is harmless to consider it); survivors are merged with the stable rule."""
    thr = best_s[:, k - 1]
    rows, cols = np.nonzero(s > thr[:, None])            # row-major: ascending column inside a row
    if rows.size == 0:
        return best_s, best_i
    starts = np.searchsorted(rows, np.arange(s.shape[0] + 1))
    for r in np.unique(rows):
        a, b = starts[r], starts[r + 1]
        cs = np.concatenate([best_s[r], s[r, cols[a:b]]])
        ci = np.concatenate([best_i[r], cols[a:b] + lo])
        order = np.argsort(-cs, kind="stable")[:k]
        best_s[r] = cs[order]
        best_i[r] = ci[order]
    return best_s, best_i


def streaming_topk(queries, candidate_batches, k, dtype=np.float64):
    """tfrs Streaming: running top-k over candidate batches by concat + top_k (earlier batch
    wins ties)."""
    cands = np.concatenate(list(candidate_batches), axis=0)
    return brute_force_topk(queries, cands, k, dtype=dtype)


def topk_merge(scores_list, ids_list, k):
    """K-way merge of per-shard top-k lists with the global rule (score desc, id asc)."""
    s = np.concatenate(scores_list, axis=1)
    i = np.concatenate(ids_list, axis=1)
    out_s = np.empty((s.shape[0], min(k, s.shape[1])), dtype=s.dtype)
    out_i = np.empty(out_s.shape, dtype=i.dtype)
    for r in range(s.shape[0]):
        order = np.lexsort((i[r], -s[r]))[:k]
        out_s[r] = s[r, order]
        out_i[r] = i[r, order]
    return out_s, out_i


# ----------------------------------------------------------------------------------------
# A.5 tfrs.metrics.FactorizedTopK
# ----------------------------------------------------------------------------------------
def factorized_topk_hits(q, true_c, topk_scores, topk_ids, ks, true_candidate_ids=None,
                         dtype=np.float64):
    """Per-row hit indicators for each k.
    Score mode: y_pred = [positive | topk_scores]; in_top_k(target=0): hit iff fewer than k
    entries are STRICTLY greater than the positive score (ties count as hits).
    Id mode: hit iff true id is among topk_ids[:, :k]."""
    hits = {}
    if true_candidate_ids is None:
        positive = np.sum(q.astype(dtype) * true_c.astype(dtype), axis=1, keepdims=True)
        y = np.concatenate([positive, topk_scores.astype(dtype)], axis=1)
        greater = (y > y[:, :1]).sum(axis=1)
        for k in ks:
            hits[k] = (greater < k).astype(np.float64)
    else:
        t = np.asarray(true_candidate_ids).reshape(-1, 1)
        for k in ks:
            hits[k] = np.clip((t == topk_ids[:, :k]).sum(axis=1), 0, 1).astype(np.float64)
    return hits


class FactorizedTopKOracle:
    """tfrs.metrics.FactorizedTopK: one weighted Mean per k, named
    factorized_top_k/top_{k}_categorical_accuracy (ks from
    /root/reference/configs/data_config.yaml:71)."""

    def __init__(self, candidates, ks=(1, 5, 10, 50, 100), identifiers=None, dtype=np.float64):
        self.candidates = candidates
        self.identifiers = identifiers
        self.ks = tuple(ks)
        self.dtype = dtype
        self.total = {k: 0.0 for k in ks}
        self.count = {k: 0.0 for k in ks}

    def update_state(self, q, true_c, true_candidate_ids=None, sample_weight=None):
        s, i = brute_force_topk(q, self.candidates, max(self.ks), self.identifiers, self.dtype)
        hits = factorized_topk_hits(q, true_c, s, i, self.ks, true_candidate_ids, self.dtype)
        w = np.ones(q.shape[0]) if sample_weight is None else np.asarray(sample_weight, dtype=np.float64)
        for k in self.ks:
            self.total[k] += float(np.sum(hits[k] * w))
            self.count[k] += float(np.sum(w))

    def result(self):
        return {f"factorized_top_k/top_{k}_categorical_accuracy":
                (self.total[k] / self.count[k] if self.count[k] else 0.0) for k in self.ks}


# ----------------------------------------------------------------------------------------
# whole tower / whole step (A.1 tfrs.models.Model.train_step)
# ----------------------------------------------------------------------------------------
class TowerSpec:
    """A tower = sum over features of (Embedding | pooled Embedding) -> optional Dense stack.
    features: list of (name, kind, vocab, mode) with kind in {"id","bag"}."""

    def __init__(self, features, dim, mlp_dims=()):
        self.features = list(features)
        self.dim = dim
        self.mlp_dims = tuple(mlp_dims)


def init_tower(spec: TowerSpec, rng: np.random.Generator, dtype=np.float32):
    p = {"tables": {}, "kernels": [], "biases": []}
    for name, _kind, vocab, _mode in spec.features:
        p["tables"][name] = keras_uniform(rng, (vocab, spec.dim), dtype)
    fan_in = spec.dim
    for u in spec.mlp_dims:
        p["kernels"].append(glorot_uniform(rng, fan_in, u, dtype))
        p["biases"].append(np.zeros(u, dtype=dtype))
        fan_in = u
    return p


def tower_forward(spec, params, inputs, dtype=np.float64, bf16=False):
    """inputs[name] = ids[B] (kind id) or (values, offsets) (kind bag)."""
    x = None
    for name, kind, _v, mode in spec.features:
        t = params["tables"][name].astype(dtype)
        if kind == "id":
            e = embedding_lookup(t, inputs[name])
        else:
            vals, offs = inputs[name]
            e = embedding_bag(t, vals, offs, mode, dtype)
        x = e if x is None else x + e
    ks = [k.astype(dtype) for k in params["kernels"]]
    bs = [b.astype(dtype) for b in params["biases"]]
    if ks:
        out, acts = mlp_forward(x, ks, bs, bf16=bf16)
    else:
        out, acts = x, []
    return out, dict(x=x, acts=acts, kernels=ks)


def tower_backward(spec, params, inputs, cache, dout):
    """Returns (dkernels, dbiases, sparse) with sparse[name] = (ids_with_duplicates, rows)."""
    if cache["kernels"]:
        dx, dk, db = mlp_backward(dout, cache["kernels"], cache["acts"] + [None])
    else:
        dx, dk, db = dout, [], []
    sparse = {}
    for name, kind, _v, mode in spec.features:
        if kind == "id":
            sparse[name] = (np.asarray(inputs[name]), dx)
        else:
            vals, offs = inputs[name]
            sparse[name] = embedding_bag_backward(vals, offs, dx, mode)
    return dk, db, sparse


def two_tower_train_step(qspec, cspec, qparams, cparams, qslots, cslots, batch_q, batch_c,
                         temperature=None, lr=0.001, eps=1e-7, l2=0.0, dtype=np.float64,
                         sample_weight=None, candidate_sampling_probability=None,
                         candidate_ids=None, remove_accidental_hits=False,
                         num_hard_negatives=None, bf16=False, inplace=False):
    """tfrs.models.Model.train_step with Keras Adagrad on every variable:
    loss = task(q, c); reg = sum(model.losses); total = loss + reg; grads of total;
    apply_gradients (sparse for tables, dense for Dense).  Params/slots are updated in place
    (as new arrays).  Returns dict(loss, regularization_loss, total_loss, q, c, dq, dc)."""
    q, qc = tower_forward(qspec, qparams, batch_q, dtype, bf16)
    c, cc = tower_forward(cspec, cparams, batch_c, dtype, bf16)
    if bf16:
        q = bf16_round(q).astype(dtype)
        c = bf16_round(c).astype(dtype)
    r = retrieval_loss_and_grads(q, c, temperature=temperature, sample_weight=sample_weight,
                                 candidate_sampling_probability=candidate_sampling_probability,
                                 candidate_ids=candidate_ids,
                                 remove_accidental_hits=remove_accidental_hits,
                                 num_hard_negatives=num_hard_negatives, dtype=dtype)
    reg = 0.0
    if l2:
        reg = l2_regularization(qparams["kernels"], l2) + l2_regularization(cparams["kernels"], l2)
    out = dict(loss=r["loss"], regularization_loss=reg, total_loss=r["loss"] + reg, q=q, c=c,
               dq=r["dq"], dc=r["dc"], unique={})
    for tag, spec, params, slots, batch, cache, dout in (
            ("q", qspec, qparams, qslots, batch_q, qc, r["dq"]),
            ("c", cspec, cparams, cslots, batch_c, cc, r["dc"])):
        dk, db, sparse = tower_backward(spec, params, batch, cache, dout)
        for l in range(len(dk)):
            g = dk[l] + (2.0 * l2 * params["kernels"][l].astype(dtype) if l2 else 0.0)
            w, a = adagrad_dense(params["kernels"][l].astype(dtype), slots["kernels"][l].astype(dtype), g, lr, eps)
            params["kernels"][l], slots["kernels"][l] = w, a
            w, a = adagrad_dense(params["biases"][l].astype(dtype), slots["biases"][l].astype(dtype), db[l], lr, eps)
            params["biases"][l], slots["biases"][l] = w, a
        for name, (ids, rows) in sparse.items():
            t, a, u = adagrad_sparse(params["tables"][name].astype(dtype, copy=False),
                                     slots["tables"][name].astype(dtype, copy=False), ids,
                                     rows.astype(dtype, copy=False), lr, eps, inplace=inplace)
            params["tables"][name], slots["tables"][name] = t, a
            out["unique"][f"{tag}/{name}"] = u
    return out
