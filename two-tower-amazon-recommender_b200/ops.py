"""Thin torch-tensor wrappers over the libtwotower C-ABI.

torch is used for device memory, the current stream and (elsewhere) torch.distributed only;
every computation below is a kernel of libtwotower.so.  Nothing here falls back to torch ops
or to the CPU: a missing library or a CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import TT_BF16, TT_F32, TT_POOL_MEAN, TT_POOL_SUM, check

# number of libtwotower kernel launches issued through this module (bench.py's gpu_launches)
LAUNCHES = 0


def _count(n: int) -> None:
    global LAUNCHES
    LAUNCHES += n


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor], dtype=None) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise TypeError("libtwotower operates on CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise ValueError("libtwotower needs contiguous tensors")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    return t.data_ptr()


def precision_code(precision: str) -> int:
    if precision == "fp32":
        return TT_F32
    if precision == "bf16":
        return TT_BF16
    raise ValueError(f"precision must be 'fp32' or 'bf16', got {precision!r}")


def pool_code(mode: str) -> int:
    if mode == "sum":
        return TT_POOL_SUM
    if mode == "mean":
        return TT_POOL_MEAN
    raise ValueError(f"combiner must be 'sum' or 'mean', got {mode!r}")


class PeerTable:
    """A table row-sharded over `world` ranks and addressed through peer-mapped base pointers (NVLink P2P):
    `ptrs` is a device int64 [world] array of the shard base addresses as seen from THIS rank (its own shard
    and the peers', e.g. from torch.distributed._symmetric_memory), `shard` this rank's [rows, d] slice.
    Row id lives on rank id % world at local row id // world."""

    def __init__(self, ptrs: torch.Tensor, shard: torch.Tensor, vocab: int, world: int, rank: int):
        self.ptrs, self.shard, self.vocab, self.world, self.rank = ptrs, shard, int(vocab), int(world), int(rank)
        self.shape = (self.vocab, shard.shape[1])
        self.device = shard.device


def _fill_feature(dst, table, values, offsets, mode) -> None:
    if isinstance(table, PeerTable) and table.world >= 2:
        dst.table = _ptr(table.ptrs, torch.int64)
        dst.vocab = table.vocab
        dst.shard_world = table.world
    elif isinstance(table, PeerTable):            # a one-rank group: the shard IS the table
        dst.table = _ptr(table.shard, torch.float32)
        dst.vocab = table.vocab
        dst.shard_world = 0
    else:
        dst.table = _ptr(table, torch.float32)
        dst.vocab = table.shape[0]
        dst.shard_world = 0
    dst.values = _ptr(values, torch.int64)
    dst.offsets = _ptr(offsets, torch.int64)
    dst.mode = pool_code(mode)


def device_check() -> None:
    check(_lib.load().tt_device_check())


# ------------------------------------------------------------------------------------ K1
def tower_input_fwd(features: Sequence[tuple], batch: int, dim: int, want_f32: bool, want_bf16: bool,
                    fault_flag: Optional[torch.Tensor] = None):
    """features: [(table f32 [V,d], values i64, offsets i64 | None, combiner)].  Returns
    (out_f32 | None, out_bf16 | None), each [batch, dim]."""
    lib = _lib.load()
    n = len(features)
    arr = (_lib.tt_feature * n)()
    dev = features[0][0].device
    for i, (table, values, offsets, mode) in enumerate(features):
        _fill_feature(arr[i], table, values, offsets, mode)
        if table.shape[1] != dim:
            raise ValueError("all features of a tower must share the embedding dimension")
    out_f32 = torch.empty((batch, dim), dtype=torch.float32, device=dev) if want_f32 else None
    out_bf16 = torch.empty((batch, dim), dtype=torch.bfloat16, device=dev) if want_bf16 else None
    if batch == 0:
        return out_f32, out_bf16
    check(lib.tt_tower_input_fwd(arr, n, _ptr(out_f32), _ptr(out_bf16), batch, dim,
                                 _ptr(fault_flag, torch.int32), _stream()))
    _count(1)
    return out_f32, out_bf16


def embedding_gather(table, ids, out_dtype=torch.float32):
    lib = _lib.load()
    out = torch.empty((ids.numel(), table.shape[1]), dtype=out_dtype, device=table.device)
    if ids.numel() == 0:
        return out
    fn = lib.tt_embedding_gather_f32 if out_dtype == torch.float32 else lib.tt_embedding_gather_bf16
    check(fn(_ptr(table, torch.float32), _ptr(ids, torch.int64), _ptr(out), ids.numel(), table.shape[1],
             table.shape[0], _stream()))
    _count(1)
    return out


def embedding_bag(table, values, offsets, mode="mean", out_dtype=torch.float32):
    lib = _lib.load()
    nb = offsets.numel() - 1
    out = torch.empty((nb, table.shape[1]), dtype=out_dtype, device=table.device)
    check(lib.tt_embedding_bag_fwd(_ptr(table, torch.float32), _ptr(values, torch.int64), _ptr(offsets, torch.int64),
                                   pool_code(mode), _ptr(out), TT_F32 if out_dtype == torch.float32 else TT_BF16,
                                   nb, table.shape[1], table.shape[0], _stream()))
    _count(1)
    return out


# ------------------------------------------------------------------------------------ K5
class SparseWorkspace:
    """Caller-owned scratch of the sparse optimizer (hash table + accumulation rows)."""

    def __init__(self, nnz: int, dim: int, device):
        lib = _lib.load()
        self.nnz, self.dim = int(nnz), int(dim)
        self.nbytes = int(lib.tt_sparse_workspace_bytes(self.nnz, self.dim))
        self.buf = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        check(lib.tt_sparse_workspace_init(_ptr(self.buf), self.nbytes, self.nnz, self.dim, _stream()))
        _count(1)

    def fits(self, nnz: int, dim: int) -> bool:
        return nnz <= self.nnz and dim == self.dim and \
            int(_lib.load().tt_sparse_workspace_bytes(nnz, dim)) <= self.nbytes


def sparse_adagrad_update(table, accum, values, offsets, mode, grad, lr, eps, ws: SparseWorkspace,
                          first_flag=None):
    lib = _lib.load()
    nnz = values.numel()
    check(lib.tt_sparse_adagrad_update(_ptr(table, torch.float32), _ptr(accum, torch.float32), table.shape[0],
                                       table.shape[1], _ptr(values, torch.int64), _ptr(offsets, torch.int64),
                                       pool_code(mode), grad.shape[0], nnz, _ptr(grad, torch.float32), lr, eps,
                                       _ptr(ws.buf), ws.nbytes, _ptr(first_flag, torch.uint8), _stream()))
    _count(3 if nnz else 0)


def sparse_lazy_adam_update(table, m, v, values, offsets, mode, grad, alpha, beta1, beta2, eps,
                            ws: SparseWorkspace, first_flag=None):
    lib = _lib.load()
    nnz = values.numel()
    check(lib.tt_sparse_lazy_adam_update(_ptr(table, torch.float32), _ptr(m, torch.float32), _ptr(v, torch.float32),
                                         table.shape[0], table.shape[1], _ptr(values, torch.int64),
                                         _ptr(offsets, torch.int64), pool_code(mode), grad.shape[0], nnz,
                                         _ptr(grad, torch.float32), alpha, beta1, beta2, eps, _ptr(ws.buf),
                                         ws.nbytes, _ptr(first_flag, torch.uint8), _stream()))
    _count(3 if nnz else 0)


def dense_adagrad_update(w, accum, grad_parts, num_parts, lr, eps, l2=0.0, shadow=None):
    lib = _lib.load()
    rows, cols = (w.shape[0], w.shape[1]) if w.dim() == 2 else (1, w.numel())
    check(lib.tt_dense_adagrad_update(_ptr(w, torch.float32), _ptr(accum, torch.float32),
                                      _ptr(grad_parts, torch.float32), num_parts, rows, cols, lr, eps, l2,
                                      _ptr(shadow, torch.bfloat16), _stream()))
    _count(1)


def dense_adam_update(w, m, v, grad_parts, num_parts, alpha, beta1, beta2, eps, l2=0.0, shadow=None):
    lib = _lib.load()
    rows, cols = (w.shape[0], w.shape[1]) if w.dim() == 2 else (1, w.numel())
    check(lib.tt_dense_adam_update(_ptr(w, torch.float32), _ptr(m, torch.float32), _ptr(v, torch.float32),
                                   _ptr(grad_parts, torch.float32), num_parts, rows, cols, alpha, beta1, beta2,
                                   eps, l2, _ptr(shadow, torch.bfloat16), _stream()))
    _count(1)


def dense_update_multi(kind: str, items, hyper):
    """One launch for all dense variables.  items: [(w, slot0, slot1 | None, grad_parts, num_parts, l2, shadow | None)];
    hyper: (lr, eps) for "adagrad", (alpha, beta1, beta2, eps) for "adam"."""
    lib = _lib.load()
    for lo in range(0, len(items), _lib.TT_MAX_DENSE_VARS):
        chunk = items[lo:lo + _lib.TT_MAX_DENSE_VARS]
        arr = (_lib.tt_dense_var * len(chunk))()
        for i, (w, s0, s1, parts, num_parts, l2, shadow) in enumerate(chunk):
            arr[i].w = _ptr(w, torch.float32)
            arr[i].slot0 = _ptr(s0, torch.float32)
            arr[i].slot1 = _ptr(s1, torch.float32)
            arr[i].grad_parts = _ptr(parts, torch.float32)
            arr[i].n = w.numel()
            arr[i].num_parts = int(num_parts)
            arr[i].l2 = float(l2)
            arr[i].shadow = _ptr(shadow, torch.bfloat16)
        if kind == "adagrad":
            check(lib.tt_dense_adagrad_update_multi(arr, len(chunk), hyper[0], hyper[1], _stream()))
        else:
            check(lib.tt_dense_adam_update_multi(arr, len(chunk), hyper[0], hyper[1], hyper[2], hyper[3], _stream()))
        _count(1)


def sparse_update_multi(kind: str, items, hyper):
    """Three launches (insert, accumulate, apply) for all tables.  items: [(table, slot0, slot1 | None, values,
    offsets | None, mode, grad, SparseWorkspace, first_flag | None)]."""
    lib = _lib.load()
    for lo in range(0, len(items), _lib.TT_MAX_SPARSE_VARS):
        chunk = items[lo:lo + _lib.TT_MAX_SPARSE_VARS]
        arr = (_lib.tt_sparse_var * len(chunk))()
        any_nnz = False
        for i, (table, s0, s1, values, offsets, mode, grad, ws, first_flag) in enumerate(chunk):
            arr[i].table = _ptr(table, torch.float32)
            arr[i].slot0 = _ptr(s0, torch.float32)
            arr[i].slot1 = _ptr(s1, torch.float32)
            arr[i].vocab, arr[i].d = table.shape
            arr[i].values = _ptr(values, torch.int64)
            arr[i].offsets = _ptr(offsets, torch.int64)
            arr[i].num_rows = grad.shape[0]
            arr[i].nnz = values.numel()
            arr[i].grad = _ptr(grad, torch.float32)
            arr[i].workspace = _ptr(ws.buf)
            arr[i].workspace_bytes = ws.nbytes
            arr[i].first_flag = _ptr(first_flag, torch.uint8)
            arr[i].mode = pool_code(mode)
            any_nnz = any_nnz or values.numel() > 0
        if kind == "adagrad":
            check(lib.tt_sparse_adagrad_update_multi(arr, len(chunk), hyper[0], hyper[1], _stream()))
        else:
            check(lib.tt_sparse_lazy_adam_update_multi(arr, len(chunk), hyper[0], hyper[1], hyper[2], hyper[3], _stream()))
        _count(3 if any_nnz else 0)


def _fill_dense_vars(items):
    arr = (_lib.tt_dense_var * max(len(items), 1))()
    for i, (w, s0, s1, parts, num_parts, l2, shadow) in enumerate(items):
        arr[i].w = _ptr(w, torch.float32)
        arr[i].slot0 = _ptr(s0, torch.float32)
        arr[i].slot1 = _ptr(s1, torch.float32)
        arr[i].grad_parts = _ptr(parts, torch.float32)
        arr[i].n = w.numel()
        arr[i].num_parts = int(num_parts)
        arr[i].l2 = float(l2)
        arr[i].shadow = _ptr(shadow, torch.bfloat16)
    return arr


def _fill_sparse_vars(items):
    """items: [(table, slot0 | None, slot1 | None, values, offsets | None, mode, grad | None, SparseWorkspace,
    first_flag | None[, (world, rank)])] -- with (world, rank) `values` are GLOBAL ids of a table row-sharded
    over the ranks and `table` is this rank's shard."""
    arr = (_lib.tt_sparse_var * max(len(items), 1))()
    for i, item in enumerate(items):
        table, s0, s1, values, offsets, mode, grad, ws, first_flag = item[:9]
        shard = item[9] if len(item) > 9 else None
        arr[i].shard = 0 if shard is None else (int(shard[0]) << 16) | int(shard[1])
        arr[i].table = _ptr(table, torch.float32)
        arr[i].slot0 = _ptr(s0, torch.float32)
        arr[i].slot1 = _ptr(s1, torch.float32)
        arr[i].vocab, arr[i].d = table.shape
        arr[i].values = _ptr(values, torch.int64)
        arr[i].offsets = _ptr(offsets, torch.int64)
        arr[i].nnz = values.numel()
        arr[i].num_rows = values.numel() if offsets is None else offsets.numel() - 1
        arr[i].grad = _ptr(grad, torch.float32)
        arr[i].workspace = _ptr(ws.buf)
        arr[i].workspace_bytes = ws.nbytes
        arr[i].first_flag = _ptr(first_flag, torch.uint8)
        arr[i].mode = pool_code(mode)
    return arr


def bucket_layout(shapes):
    """Offsets (in floats, each variable padded to a multiple of 4) of the flat dense-gradient bucket."""
    offs, off = [], 0
    for shp in shapes:
        offs.append(off)
        off += (int(torch.Size(shp).numel()) + 3) // 4 * 4
    return offs, off


def fold_parts_into_bucket(grads, out=None):
    """grads: [(parts [P, ...] f32, P)].  Returns (bucket f32 [sum n_i], [views]) with view_i = ordered sum of parts_i,
    all in one launch (the flat bucket is what the data-parallel all-reduce sends).  out: caller-owned bucket."""
    lib = _lib.load()
    sizes = [p[0].numel() for p, _ in grads]
    padded = [(n + 3) // 4 * 4 for n in sizes]
    bucket = out if out is not None else torch.empty(sum(padded), dtype=torch.float32, device=grads[0][0].device)
    if bucket.numel() < sum(padded):
        raise ValueError("fold_parts_into_bucket: destination bucket too small")
    views, off = [], 0
    for (p, _), n, m in zip(grads, sizes, padded):
        views.append(bucket[off:off + n].view(p[0].shape))
        off += m
    for lo in range(0, len(grads), _lib.TT_MAX_DENSE_VARS):
        chunk = list(zip(grads, views))[lo:lo + _lib.TT_MAX_DENSE_VARS]
        arr = (_lib.tt_dense_var * len(chunk))()
        for i, ((p, P), v) in enumerate(chunk):
            arr[i].w = v.data_ptr()
            arr[i].grad_parts = _ptr(p, torch.float32)
            arr[i].n = v.numel()
            arr[i].num_parts = int(P)
        check(lib.tt_fold_parts_multi(arr, len(chunk), _stream()))
        _count(1)
    return bucket, views


def sparse_prepare(items):
    """Hash insert of the ids of every table (one launch): needs no gradients, so it can run on a side
    stream while the forward pass executes.  items as in _fill_sparse_vars (slots / grad may be None)."""
    if not items:
        return
    check(_lib.load().tt_optimizer_prepare_sparse(_fill_sparse_vars(items), len(items), _stream()))
    _count(1 if any(it[3].numel() for it in items) else 0)


def adam_bias_correction(step_dev: torch.Tensor, alpha_dev: torch.Tensor, lr: float, beta1: float, beta2: float) -> None:
    """t = ++step_dev[0]; alpha_dev[0] = lr * sqrt(1 - beta2^t) / (1 - beta1^t), on the device (graph-replayable)."""
    check(_lib.load().tt_adam_bias_correction(_ptr(step_dev, torch.int64), lr, beta1, beta2, _ptr(alpha_dev, torch.float32), _stream()))
    _count(1)


def optimizer_step(kind: str, dense_items, sparse_items, hyper):
    """ONE launch: every dense variable and every (prepared) table.  kind "adagrad": hyper = (lr, eps);
    "lazy_adam": hyper = (alpha float | device fp32 [1] tensor, beta1, beta2, eps)."""
    lib = _lib.load()
    if len(dense_items) > _lib.TT_MAX_DENSE_VARS or len(sparse_items) > _lib.TT_MAX_SPARSE_VARS:
        raise ValueError("optimizer_step: too many variables for one launch; split the call")
    d, s = _fill_dense_vars(dense_items), _fill_sparse_vars(sparse_items)
    if kind == "adagrad":
        check(lib.tt_adagrad_step(d, len(dense_items), s, len(sparse_items), hyper[0], hyper[1], _stream()))
    else:
        a = hyper[0]
        a_host, a_dev = (0.0, _ptr(a, torch.float32)) if isinstance(a, torch.Tensor) else (float(a), None)
        check(lib.tt_lazy_adam_step(d, len(dense_items), s, len(sparse_items), a_host, a_dev, hyper[1], hyper[2], hyper[3], _stream()))
    _count(1)


def sum_squares(x, scale, out, accumulate):
    check(_lib.load().tt_sum_squares(_ptr(x, torch.float32), x.numel(), scale, _ptr(out, torch.float32),
                                     1 if accumulate else 0, _stream()))
    _count(1)


# ------------------------------------------------------------------------------------ K2
def dense_fwd(precision: str, x, kernel, bias, relu: bool, want_f32: bool = False):
    """y = act(x @ kernel + bias), kernel in the Keras layout [in, out].
    fp32: fp32 tensors -> (y f32, None).  bf16: x bf16, kernel bf16 (shadow) -> (y bf16, y_f32 | None)."""
    lib = _lib.load()
    M, in_dim = x.shape
    out_dim = kernel.shape[1]
    if precision == "fp32":
        y = torch.empty((M, out_dim), dtype=torch.float32, device=x.device)
        check(lib.tt_dense_fwd(TT_F32, _ptr(x, torch.float32), _ptr(kernel, torch.float32),
                               _ptr(bias, torch.float32), _ptr(y), None, M, in_dim, out_dim,
                               1 if relu else 0, _stream()))
        _count(1)
        return y, None
    y = torch.empty((M, out_dim), dtype=torch.bfloat16, device=x.device)
    y_f32 = torch.empty((M, out_dim), dtype=torch.float32, device=x.device) if want_f32 else None
    check(lib.tt_dense_fwd(TT_BF16, _ptr(x, torch.bfloat16), _ptr(kernel, torch.bfloat16),
                           _ptr(bias, torch.float32), _ptr(y), _ptr(y_f32), M, in_dim, out_dim,
                           1 if relu else 0, _stream()))
    _count(1)
    return y, y_f32


def dense_bwd_num_parts(precision: str, M: int, in_dim: int, out_dim: int) -> int:
    return int(_lib.load().tt_dense_bwd_num_parts(precision_code(precision), M, in_dim, out_dim))


def colsum_f32(x, num_parts: int = 0):
    """Column sums as `num_parts` stacked partial rows [P, cols] (summed later, in order, by the optimizer)."""
    rows, cols = x.shape
    P = num_parts or max(1, min(64, rows // 64))
    out = torch.empty((P, cols), dtype=torch.float32, device=x.device)
    check(_lib.load().tt_colsum_f32(_ptr(x, torch.float32), _ptr(out), rows, cols, P, _stream()))
    _count(1)
    return out


def sum_parts(parts, num_parts: int):
    """[P, ...] -> [...] (fixed order)."""
    n = parts[0].numel()
    out = torch.empty_like(parts[0])
    check(_lib.load().tt_sum_parts_f32(_ptr(parts, torch.float32), num_parts, n, _ptr(out), _stream()))
    _count(1)
    return out


def dense_bwd(precision: str, dy, x, kernel, relu_mask_x: bool, want_dx: bool, want_dx_f32: bool = False,
              want_dbias: bool = True):
    """Returns (dx, dx_f32, dkernel_parts [P,in,out] f32, P, dbias_parts [P,out] f32 | None)."""
    lib = _lib.load()
    M, out_dim = dy.shape
    in_dim = x.shape[1]
    dev = dy.device
    P = dense_bwd_num_parts(precision, M, in_dim, out_dim)
    dk = torch.empty((P, in_dim, out_dim), dtype=torch.float32, device=dev)
    db = torch.empty((P, out_dim), dtype=torch.float32, device=dev) if want_dbias else None
    if precision == "fp32":
        dx = torch.empty((M, in_dim), dtype=torch.float32, device=dev) if want_dx else None
        check(lib.tt_dense_bwd(TT_F32, _ptr(dy, torch.float32), _ptr(x, torch.float32), _ptr(kernel, torch.float32),
                               _ptr(dx), None, _ptr(dk), P, _ptr(db), M, in_dim, out_dim,
                               1 if relu_mask_x else 0, _stream()))
        _count(1 + int(want_dx) + int(want_dbias))
        return dx, None, dk, P, db
    dx = torch.empty((M, in_dim), dtype=torch.bfloat16, device=dev) if want_dx else None
    dx_f32 = torch.empty((M, in_dim), dtype=torch.float32, device=dev) if want_dx_f32 else None
    check(lib.tt_dense_bwd(TT_BF16, _ptr(dy, torch.bfloat16), _ptr(x, torch.bfloat16), _ptr(kernel, torch.bfloat16),
                           _ptr(dx), _ptr(dx_f32), _ptr(dk), P, _ptr(db), M, in_dim, out_dim,
                           1 if relu_mask_x else 0, _stream()))
    _count(1 + int(want_dx or want_dx_f32) + int(want_dbias))
    return dx, dx_f32, dk, P, db


# ------------------------------------------------------------------------ K1+K2 fused tower
def tower_mlp2_supported(d_in: int, d_hid: int, d_out: int) -> bool:
    return bool(_lib.load().tt_tower_mlp2_supported(int(d_in), int(d_hid), int(d_out)))


def _fill_tower(dst, t):
    feats = t["features"]
    dst.num_feats = len(feats)
    for i, (table, values, offsets, mode) in enumerate(feats):
        _fill_feature(dst.feats[i], table, values, offsets, mode)
    dst.d_in, dst.d_hid = t["w1"].shape
    dst.d_out = t["w2"].shape[1]
    dst.batch = t["batch"]
    dst.w1, dst.b1 = _ptr(t["w1"], torch.bfloat16), _ptr(t["b1"], torch.float32)
    dst.w2, dst.b2 = _ptr(t["w2"], torch.bfloat16), _ptr(t["b2"], torch.float32)
    ws = t.get("prepare_ws")                    # SparseWorkspace: the forward kernel also dedups this tower's ids
    dst.prepare_workspace = _ptr(ws.buf) if ws is not None else None
    dst.prepare_workspace_bytes = ws.nbytes if ws is not None else 0


def tower_mlp2_fwd(towers, fault_flag: Optional[torch.Tensor] = None):
    """towers: [dict(features=[(table, values, offsets, mode)], batch, w1 bf16 [in,hid], b1, w2 bf16 [hid,out], b2
    [, prepare_ws=SparseWorkspace])].  One launch for all towers.  With prepare_ws (ID-only towers) the kernel also
    performs sparse_prepare for the tower's ids.  Returns [(x bf16 [B,in], h bf16 [B,hid], y bf16 [B,out])]."""
    lib = _lib.load()
    arr = (_lib.tt_tower_mlp2 * len(towers))()
    outs = []
    for i, t in enumerate(towers):
        _fill_tower(arr[i], t)
        dev, B = t["w1"].device, t["batch"]
        x = t.get("x_input")                      # given tower input (row-sharded lookup): features must be []
        if x is None:
            x = torch.empty((B, arr[i].d_in), dtype=torch.bfloat16, device=dev)
        elif t["features"]:
            raise ValueError("tower_mlp2_fwd: pass either features or x_input")
        h = torch.empty((B, arr[i].d_hid), dtype=torch.bfloat16, device=dev)
        y = torch.empty((B, arr[i].d_out), dtype=torch.bfloat16, device=dev)
        arr[i].x, arr[i].h, arr[i].y = _ptr(x, torch.bfloat16), _ptr(h), _ptr(y)
        outs.append((x, h, y))
    check(lib.tt_tower_mlp2_fwd(arr, len(towers), _ptr(fault_flag, torch.int32), _stream()))
    _count(1)
    return outs


def tower_mlp2_bwd(towers):
    """towers: the forward dicts plus x, h (saved) and dy_parts fp32 [S, B, out], dy_splits.  One launch.
    Returns [dict(dx f32 [B,in], dw1 [P,in,hid], dw2 [P,hid,out], db1 [P,hid], db2 [P,out], P)]."""
    lib = _lib.load()
    arr = (_lib.tt_tower_mlp2 * len(towers))()
    outs = []
    for i, t in enumerate(towers):
        _fill_tower(arr[i], dict(t, features=[], prepare_ws=None))
        dev, B = t["w1"].device, t["batch"]
        d_in, d_hid, d_out = arr[i].d_in, arr[i].d_hid, arr[i].d_out
        P = (B + 127) // 128
        o = dict(dx=torch.empty((B, d_in), dtype=torch.float32, device=dev),
                 dw1=torch.empty((P, d_in, d_hid), dtype=torch.float32, device=dev),
                 dw2=torch.empty((P, d_hid, d_out), dtype=torch.float32, device=dev),
                 db1=torch.empty((P, d_hid), dtype=torch.float32, device=dev),
                 db2=torch.empty((P, d_out), dtype=torch.float32, device=dev), P=P)
        arr[i].x, arr[i].h = _ptr(t["x"], torch.bfloat16), _ptr(t["h"], torch.bfloat16)
        arr[i].dy_parts, arr[i].dy_splits = _ptr(t["dy_parts"], torch.float32), int(t["dy_splits"])
        arr[i].dx, arr[i].dw1_parts, arr[i].dw2_parts = _ptr(o["dx"]), _ptr(o["dw1"]), _ptr(o["dw2"])
        arr[i].db1_parts, arr[i].db2_parts = _ptr(o["db1"]), _ptr(o["db2"])
        outs.append(o)
    check(lib.tt_tower_mlp2_bwd(arr, len(towers), _stream()))
    _count(1)
    return outs


def cast_f32_to_bf16(x, out=None):
    """out: existing bf16 tensor of the same shape to cast INTO (keeps its address: CUDA graphs, shadows)."""
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    elif out.shape != x.shape or out.dtype != torch.bfloat16:
        raise ValueError("cast_f32_to_bf16: `out` must be a bf16 tensor of the input's shape")
    check(_lib.load().tt_cast_f32_to_bf16(_ptr(x, torch.float32), _ptr(out), x.numel(), _stream()))
    _count(1)
    return out


# --------------------------------------------------------------------------------- K3/K4
_ws_cache = {}
_ws_retired = []      # outgrown scratch buffers: kept alive, because a captured CUDA graph may still launch kernels on them


def _workspace(nbytes: int, device, tag: str = "scratch") -> torch.Tensor:
    """Scratch reused across calls on one device (grown on demand, geometrically), one buffer per user (`tag`).
    Buffers are zero-filled when allocated: the retrieval forward keeps arrival tickets at the head of its workspace
    (tt_retrieval_workspace_init contract) and leaves them zero after every launch.  A buffer that has to grow is
    RETIRED, not freed: a CUDA graph captured earlier keeps replaying on the old one (each buffer is self-contained)."""
    key = (device.type, device.index, tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _ws_retired.append(buf)
        grown = 0 if buf is None else 2 * buf.numel()
        buf = torch.zeros(max(nbytes, grown, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def retrieval_loss_fwd(precision: str, q, c, inv_temperature: float, label_offset: int = 0, sample_weight=None,
                       cand_log_q=None, cand_ids=None):
    """Returns (loss [1] f32, row_lse [nq] f32, row_pos [nq] f32)."""
    lib = _lib.load()
    pc = precision_code(precision)
    dt = torch.float32 if precision == "fp32" else torch.bfloat16
    nq, d = q.shape
    nc = c.shape[0]
    dev = q.device
    lse = torch.empty((nq,), dtype=torch.float32, device=dev)
    pos = torch.empty((nq,), dtype=torch.float32, device=dev)
    loss = _new_loss(dev)
    nbytes = int(lib.tt_retrieval_workspace_bytes(pc, nq, nc, d))
    ws = _workspace(nbytes, dev, "retrieval")
    check(lib.tt_retrieval_loss_fwd(pc, _ptr(q, dt), _ptr(c, dt), nq, nc, d, inv_temperature, label_offset,
                                    _ptr(sample_weight, torch.float32), _ptr(cand_log_q, torch.float32),
                                    _ptr(cand_ids, torch.int64), _ptr(lse), _ptr(pos), _ptr(loss), _ptr(ws),
                                    ws.numel(), _stream()))
    _count(1 if precision == "bf16" else 2)
    return loss, lse, pos


def retrieval_loss_bwd(precision: str, q, c, inv_temperature: float, row_lse, label_offset: int = 0,
                       sample_weight=None, cand_log_q=None, cand_ids=None, grad_scale: float = 1.0,
                       want_bf16=(False, False)):
    """Returns dict(dq, dc [f32], dq_bf16, dc_bf16)."""
    lib = _lib.load()
    pc = precision_code(precision)
    dt = torch.float32 if precision == "fp32" else torch.bfloat16
    nq, d = q.shape
    nc = c.shape[0]
    dev = q.device
    dq = torch.empty((nq, d), dtype=torch.float32, device=dev)
    dc = torch.empty((nc, d), dtype=torch.float32, device=dev)
    mk = lambda shape, on: torch.empty(shape, dtype=torch.bfloat16, device=dev) if on else None
    dq_b, dc_b = mk((nq, d), want_bf16[0]), mk((nc, d), want_bf16[1])
    nbytes = int(lib.tt_retrieval_workspace_bytes(pc, nq, nc, d))
    ws = _workspace(nbytes, dev, "retrieval")
    check(lib.tt_retrieval_loss_bwd(pc, _ptr(q, dt), _ptr(c, dt), nq, nc, d, inv_temperature,
                                    label_offset, _ptr(sample_weight, torch.float32),
                                    _ptr(cand_log_q, torch.float32), _ptr(cand_ids, torch.int64),
                                    _ptr(row_lse, torch.float32), grad_scale, _ptr(dq), _ptr(dc), _ptr(dq_b),
                                    _ptr(dc_b), _ptr(ws), ws.numel(), _stream()))
    _count(4 if precision == "bf16" else 2)
    return dict(dq=dq, dc=dc, dq_bf16=dq_b, dc_bf16=dc_b)


def retrieval_bwd_num_splits(nq: int, nc: int, d: int):
    sq, sc = C.c_int32(0), C.c_int32(0)
    check(_lib.load().tt_retrieval_bwd_num_splits(TT_BF16, nq, nc, d, C.byref(sq), C.byref(sc)))
    return sq.value, sc.value


# Where the scalar loss of a retrieval forward goes.  GraphedStep(steps_per_execution = S) sets this while it captures
# so that the S losses of one execution are the S elements of ONE device tensor (one D2H copy reads them all).
loss_allocator = None


def _new_loss(dev) -> torch.Tensor:
    if loss_allocator is not None:
        t = loss_allocator()
        if t is not None:
            return t
    return torch.empty((1,), dtype=torch.float32, device=dev)


def retrieval_fwd_dq_supported(nq: int, nc: int, d: int) -> bool:
    return int(_lib.load().tt_retrieval_fwd_dq_workspace_bytes(nq, nc, d)) > 0


_fin_stream = None
_fin_pending = False


def join_side_work() -> None:
    """Make the current stream wait for the forked fold of retrieval_loss_fwd_dq(fork=True), if one is in flight."""
    global _fin_pending
    if _fin_pending:
        torch.cuda.current_stream().wait_stream(_fin_stream)
        _fin_pending = False


def retrieval_loss_fwd_dq(q, c, inv_temperature: float, label_offset: int = 0, sample_weight=None, fork: bool = False):
    """bf16 only: the loss forward and dQ in one pass.  Returns (loss [1], row_lse [nq], row_pos [nq], dq f32 [nq, d],
    workspace).  fork=True: the last summation of the scalar loss runs on a side stream (off the critical path of the
    backward) -- call join_side_work() before anything reads `loss`."""
    global _fin_stream, _fin_pending
    lib = _lib.load()
    nq, d = q.shape
    nc = c.shape[0]
    dev = q.device
    lse = torch.empty((nq,), dtype=torch.float32, device=dev)
    pos = torch.empty((nq,), dtype=torch.float32, device=dev)
    loss = _new_loss(dev)
    dq = torch.empty((nq, d), dtype=torch.float32, device=dev)
    nbytes = int(lib.tt_retrieval_fwd_dq_workspace_bytes(nq, nc, d))
    if nbytes <= 0:
        raise ValueError(f"retrieval_loss_fwd_dq: shape not supported (d={d})")
    ws = _workspace(nbytes, dev, "retrieval_fwd_dq")
    fin = None
    if fork:
        join_side_work()
        if _fin_stream is None:
            _fin_stream = torch.cuda.Stream()
        fin = _fin_stream.cuda_stream
    check(lib.tt_retrieval_loss_fwd_dq(_ptr(q, torch.bfloat16), _ptr(c, torch.bfloat16), nq, nc, d, inv_temperature,
                                       label_offset, _ptr(sample_weight, torch.float32), _ptr(lse), _ptr(pos), _ptr(loss),
                                       _ptr(dq), _ptr(ws), ws.numel(), _stream(), fin))
    _fin_pending = fork
    _count(3)
    return loss, lse, pos, dq, ws


def retrieval_loss_bwd_dc_fused(q, c, inv_temperature: float, fwd_ws, label_offset: int = 0, sample_weight=None,
                                grad_scale: float = 1.0):
    """The dC pass after retrieval_loss_fwd_dq (column lse from its workspace).  Returns dc_parts f32 [sc, nc, d]."""
    nq, d = q.shape
    nc = c.shape[0]
    _sq, sc = retrieval_bwd_num_splits(nq, nc, d)
    dc_parts = torch.empty((sc, nc, d), dtype=torch.float32, device=q.device)
    check(_lib.load().tt_retrieval_loss_bwd_dc_fused(_ptr(q, torch.bfloat16), _ptr(c, torch.bfloat16), nq, nc, d,
                                                     inv_temperature, label_offset, _ptr(sample_weight, torch.float32),
                                                     _ptr(fwd_ws), grad_scale, _ptr(dc_parts), _stream()))
    _count(1)
    return dc_parts


def retrieval_loss_bwd_parts(q, c, inv_temperature: float, row_lse, label_offset: int = 0, sample_weight=None,
                             cand_log_q=None, cand_ids=None, grad_scale: float = 1.0, want_dq: bool = True):
    """bf16 only.  Returns (dq_parts f32 [sq, nq, d], dc_parts f32 [sc, nc, d]): dq = dq_parts.sum(0) in index order.
    want_dq=False: the dC pass only (dq came from retrieval_loss_fwd_dq); dq_parts is None."""
    nq, d = q.shape
    nc = c.shape[0]
    sq, sc = retrieval_bwd_num_splits(nq, nc, d)
    dq_parts = torch.empty((sq, nq, d), dtype=torch.float32, device=q.device) if want_dq else None
    dc_parts = torch.empty((sc, nc, d), dtype=torch.float32, device=q.device)
    check(_lib.load().tt_retrieval_loss_bwd_parts(TT_BF16, _ptr(q, torch.bfloat16), _ptr(c, torch.bfloat16), nq, nc, d,
                                                  inv_temperature, label_offset, _ptr(sample_weight, torch.float32),
                                                  _ptr(cand_log_q, torch.float32), _ptr(cand_ids, torch.int64),
                                                  _ptr(row_lse, torch.float32), grad_scale, _ptr(dq_parts), _ptr(dc_parts),
                                                  _stream()))
    _count(2 if want_dq else 1)
    return dq_parts, dc_parts


def combine_parts(parts, want_f32: bool = True, want_bf16: bool = False, out_f32=None):
    """[S, rows, d] fp32 -> ordered sum as fp32 and/or bf16 [rows, d] (out_f32: caller-owned fp32 destination)."""
    S, rows, d = parts.shape
    out_f = out_f32 if out_f32 is not None else (
        torch.empty((rows, d), dtype=torch.float32, device=parts.device) if want_f32 else None)
    out_b = torch.empty((rows, d), dtype=torch.bfloat16, device=parts.device) if want_bf16 else None
    check(_lib.load().tt_combine_parts_f32(_ptr(parts, torch.float32), S, rows, d, _ptr(out_f), _ptr(out_b), _stream()))
    _count(1)
    return out_f, out_b


# ------------------------------------------------------------------------------------ K6
def topk_bruteforce(precision: str, queries, candidates, k: int, cand_index_base: int = 0, identifiers=None,
                    uncertain: Optional[torch.Tensor] = None):
    """Exact top-k: scoring stage (k + margin pool) + exact re-rank.  uncertain: optional device int32 [1] counter of
    rows whose margin could not be shown wide enough (see include/twotower.h)."""
    lib = _lib.load()
    pc = precision_code(precision)
    dt = torch.float32 if precision == "fp32" else torch.bfloat16
    nq, d = queries.shape
    nc = candidates.shape[0]
    dev = queries.device
    scores = torch.empty((nq, k), dtype=torch.float32, device=dev)
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    if nq == 0:
        return scores, ids
    nbytes = int(lib.tt_topk_workspace_bytes(pc, nq, nc, d, k))
    ws = _workspace(nbytes, dev)
    launches = int(lib.tt_topk_num_launches(pc, nq, nc, d, k))
    check(lib.tt_topk_bruteforce(pc, _ptr(queries, dt), _ptr(candidates, dt), nq, nc, d, k, cand_index_base,
                                 _ptr(identifiers, torch.int64), _ptr(scores), _ptr(ids), _ptr(uncertain, torch.int32),
                                 _ptr(ws), ws.numel(), _stream()))
    _count(launches)
    return scores, ids


def topk_bruteforce_peer(precision: str, queries, candidates, k: int, cand_index_base: int, ws: "PeerWorkspace",
                         queries_per_rank: int, recv_scores_offset: int, recv_ids_offset: int,
                         uncertain: Optional[torch.Tensor] = None) -> None:
    """Candidate-sharded serving: score ALL queries against this rank's shard; the exact partial list of query qi lands in
    list slot [rank] of the receive area ([world, queries_per_rank, k] scores / global indices at the given offsets of
    the symmetric workspace) of rank qi // queries_per_rank.  Follow with peer_barrier + topk_merge."""
    lib = _lib.load()
    pc = precision_code(precision)
    dt = torch.float32 if precision == "fp32" else torch.bfloat16
    nq, d = queries.shape
    nc = candidates.shape[0]
    nbytes = int(lib.tt_topk_workspace_bytes(pc, nq, nc, d, k))
    scratch = _workspace(nbytes, queries.device)
    launches = int(lib.tt_topk_num_launches(pc, nq, nc, d, k))
    check(lib.tt_topk_bruteforce_peer(pc, _ptr(queries, dt), _ptr(candidates, dt), nq, nc, d, k, int(cand_index_base),
                                      _ptr(ws.bases, torch.int64), ws.world, ws.rank, int(queries_per_rank),
                                      int(recv_scores_offset), int(recv_ids_offset), _ptr(uncertain, torch.int32),
                                      _ptr(scratch), scratch.numel(), _stream()))
    _count(launches)


def topk_merge(scores, ids, k_out: int, index_base: int = 0, identifiers=None):
    """scores/ids (candidate indices): [L, nq, k_in] -> ([nq, k_out], [nq, k_out])."""
    L, nq, k_in = scores.shape
    out_s = torch.empty((nq, k_out), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((nq, k_out), dtype=torch.int64, device=scores.device)
    check(_lib.load().tt_topk_merge(_ptr(scores, torch.float32), _ptr(ids, torch.int64), L, nq, k_in, k_out,
                                    index_base, _ptr(identifiers, torch.int64), _ptr(out_s), _ptr(out_i), _stream()))
    _count(1)
    return out_s, out_i


def topk_hits(positive, topk_scores, topk_ids, true_ids, sample_weight, ks, hits_out, weight_out):
    nq, k = (topk_scores if topk_scores is not None else topk_ids).shape
    arr = (C.c_int32 * len(ks))(*[int(x) for x in ks])
    check(_lib.load().tt_topk_hits(_ptr(positive, torch.float32), _ptr(topk_scores, torch.float32),
                                   _ptr(topk_ids, torch.int64), _ptr(true_ids, torch.int64),
                                   _ptr(sample_weight, torch.float32), nq, k, arr, len(ks),
                                   _ptr(hits_out, torch.float32), _ptr(weight_out, torch.float32), _stream()))
    _count(1)


def rowwise_dot(precision: str, q, c):
    out = torch.empty((q.shape[0],), dtype=torch.float32, device=q.device)
    check(_lib.load().tt_rowwise_dot(precision_code(precision), _ptr(q), _ptr(c), _ptr(out), q.shape[0], q.shape[1],
                                     _stream()))
    _count(1)
    return out


# ------------------------------------------------------------------------ sharding helpers
def partition_ids(ids, world: int, capacity: int = 0, overflow_flag=None):
    """Stable partition by owner = id % world.  Returns (send_local_rows, perm [n], counts [world]);
    send_local_rows is [n] packed (capacity == 0) or [world * capacity] padded with -1."""
    n = ids.numel()
    send = torch.empty((world * capacity if capacity else n,), dtype=torch.int64, device=ids.device)
    perm = torch.empty((n,), dtype=torch.int64, device=ids.device)
    counts = torch.empty((world,), dtype=torch.int64, device=ids.device)
    check(_lib.load().tt_partition_ids(_ptr(ids, torch.int64), n, world, capacity, _ptr(send), _ptr(perm),
                                       _ptr(counts), _ptr(overflow_flag, torch.int32), _stream()))
    _count(1)
    return send, perm, counts


def permute_rows(x, perm, inverse: bool, out_rows: int = 0, zero_fill: bool = False):
    """inverse=False: out[perm[j]] = x[j] (out has out_rows rows); inverse=True: out[j] = x[perm[j]]."""
    n = perm.numel()
    rows = out_rows or n
    shape = (rows, x.shape[1])
    out = torch.zeros(shape, dtype=x.dtype, device=x.device) if zero_fill else torch.empty(shape, dtype=x.dtype, device=x.device)
    check(_lib.load().tt_permute_rows(_ptr(x), _ptr(perm, torch.int64), _ptr(out), n, x.shape[1] * x.element_size(),
                                      1 if inverse else 0, _stream()))
    _count(1)
    return out


# ------------------------------------------------------------------ NVLink peer-memory exchange
class PeerWorkspace:
    """Symmetric workspace of the row-sharded step (include/twotower.h, csrc/peer.cu): `local` is this rank's
    uint8 buffer (first 1 KB = barrier flags), `bases` a device int64 [world] array of every rank's
    peer-mapped base address, `step` this rank's device step counter."""

    FLAG_BYTES = 1024

    def __init__(self, local: torch.Tensor, bases: torch.Tensor, world: int, rank: int):
        self.local, self.bases, self.world, self.rank = local, bases, int(world), int(rank)
        self.step = torch.zeros(16, dtype=torch.int64, device=local.device)    # [0, 8) slot epochs, [8, 16) blocks done
        self.step[:8] = 1
        self.bases_self = bases[rank:rank + 1].repeat(world).contiguous()      # every source = my own copy (slab sums)
        # flag blocks sit at offset 0 of every copy, so the flag base table is the base table itself

    def view(self, offset: int, shape, dtype) -> torch.Tensor:
        n = int(torch.Size(shape).numel()) * torch.empty((), dtype=dtype).element_size()
        return self.local[offset:offset + n].view(dtype).view(shape)


def peer_barrier(ws: PeerWorkspace, slot: int) -> None:
    check(_lib.load().tt_peer_barrier(_ptr(ws.bases, torch.int64), _ptr(ws.step, torch.int64), ws.world, ws.rank, slot,
                                      _stream()))
    _count(1)


def peer_push(ws: PeerWorkspace, segments) -> None:
    """segments: [(src tensor, dst_offset_bytes)] (<= 4): src -> the same offset of EVERY rank's workspace."""
    n = len(segments)
    src = (C.c_void_p * n)(*[_ptr(t) for t, _ in segments])
    off = (C.c_int64 * n)(*[int(o) for _, o in segments])
    nb = (C.c_int64 * n)(*[t.numel() * t.element_size() for t, _ in segments])
    check(_lib.load().tt_peer_push(_ptr(ws.bases, torch.int64), ws.world, n, src, off, nb, _stream()))
    _count(1)


def peer_sum(ws: PeerWorkspace, offset: int, out: torch.Tensor, slot: int = -1, local_stride: int = 0) -> torch.Tensor:
    """out (fp32, numel % 4 == 0) = sum over r, in rank order, of the fp32 block at `offset` of rank r's workspace, or
    (local_stride > 0) of the `world` slabs at offset + r * local_stride of MY workspace (filled by the peers)."""
    src = ws.bases_self if local_stride else ws.bases
    check(_lib.load().tt_peer_sum_f32(_ptr(src, torch.int64), int(offset), int(local_stride), out.numel(), _ptr(out, torch.float32),
                                      _ptr(ws.bases, torch.int64), _ptr(ws.step, torch.int64), ws.world, ws.rank, slot,
                                      _stream()))
    _count(1)
    return out


def peer_pull_rows(ws: PeerWorkspace, tables, rows_per_rank: int, d: int, slot: int = -1) -> None:
    """tables: [(ids_all int64 [world*b], src_offset_bytes, out fp32 [world*b, d])] (<= 4)."""
    n = len(tables)
    ids = (C.c_void_p * n)(*[_ptr(i, torch.int64) for i, _, _ in tables])
    off = (C.c_int64 * n)(*[int(o) for _, o, _ in tables])
    out = (C.c_void_p * n)(*[_ptr(o, torch.float32) for _, _, o in tables])
    check(_lib.load().tt_peer_pull_rows(_ptr(ws.bases, torch.int64), n, ids, off, out, int(rows_per_rank), int(d),
                                        _ptr(ws.bases, torch.int64), _ptr(ws.step, torch.int64), ws.world, ws.rank, slot,
                                        _stream()))
    _count(1)


def peer_combine_scatter(ws: PeerWorkspace, parts: torch.Tensor, rows_per_rank: int, dst_offset: int) -> None:
    """Ordered sum of parts [S, world * b, d]; row i lands in slot [rank] of the receive area of rank i // b."""
    S, rows, d = parts.shape
    check(_lib.load().tt_peer_combine_scatter(_ptr(ws.bases, torch.int64), _ptr(parts, torch.float32), S, rows, d,
                                              int(rows_per_rank), int(dst_offset), ws.world, ws.rank, _stream()))
    _count(1)


def peer_push_rows(ws: PeerWorkspace, tables, b: int, d: int) -> None:
    """tables: [(ids int64 [b], rows fp32 [b, d], dst_offset_bytes)] (<= 4): row j -> the owner of table row ids[j]."""
    n = len(tables)
    ids = (C.c_void_p * n)(*[_ptr(i, torch.int64) for i, _, _ in tables])
    src = (C.c_void_p * n)(*[_ptr(r, torch.float32) for _, r, _ in tables])
    off = (C.c_int64 * n)(*[int(o) for _, _, o in tables])
    check(_lib.load().tt_peer_push_rows(_ptr(ws.bases, torch.int64), n, ids, src, off, int(b), int(d), ws.world, ws.rank, _stream()))
    _count(1)


def peer_row_maps(peer_addrs, rows: int, d: int, device) -> torch.Tensor:
    """Device array [world] of TMA tensor maps over every rank's [rows, d] fp32 area (peer-mapped addresses)."""
    n = len(peer_addrs)
    addrs = (C.c_uint64 * n)(*[int(a) for a in peer_addrs])
    host = torch.zeros(n * 128, dtype=torch.uint8)
    check(_lib.load().tt_peer_make_row_maps(addrs, n, int(rows), int(d), host.data_ptr()))
    dev = torch.empty(n * 128 + 64, dtype=torch.uint8, device=device)
    off = (-dev.data_ptr()) % 64                      # tensor maps must be 64-byte aligned
    out = dev[off:off + n * 128]
    out.copy_(host)
    torch.cuda.synchronize()
    return out


def peer_retrieval_bwd_dc(ws: PeerWorkspace, maps: torch.Tensor, q, c, inv_temperature: float, row_lse, label_offset: int,
                          sample_weight, scratch: torch.Tensor) -> None:
    """The dC pass with the reduce-scatter in its epilogue: every 128-row block of dC is TMA-stored into slot [rank] of
    the receive area of the rank owning those candidates."""
    nq, d = q.shape
    nc = c.shape[0]
    check(_lib.load().tt_peer_retrieval_bwd_dc(_ptr(q, torch.bfloat16), _ptr(c, torch.bfloat16), nq, nc, d, inv_temperature,
                                               label_offset, _ptr(sample_weight, torch.float32), _ptr(row_lse, torch.float32),
                                               1.0, _ptr(maps), ws.world, ws.rank, _ptr(scratch, torch.float32), _stream()))
    _count(1)


# ------------------------------------------------------------------ hard-negative mining (tfrs num_hard_negatives)
def select_hard_negatives(precision: str, q, c, num_hard_negatives: int, label_offset: int = 0) -> torch.Tensor:
    """[nq, k] int64, k = min(n + 1, nc): column 0 = the positive (label_offset + row), then the n highest-scoring other
    candidates in tf.math.top_k order (score desc, index asc) -- tfrs HardNegativeMining's top_k(scores + labels * MAX).
    The scores come from the brute-force top-k kernel; the rest is index bookkeeping."""
    nq, nc = q.shape[0], c.shape[0]
    k = min(int(num_hard_negatives) + 1, nc)
    _scores, ids = topk_bruteforce(precision, q, c, k)
    label = torch.arange(nq, device=q.device, dtype=torch.int64) + int(label_offset)
    is_pos = ids == label[:, None]
    order = torch.argsort(is_pos.to(torch.int8), dim=1, stable=True)          # negatives first, original order kept
    neg = torch.gather(ids, 1, order)[:, :k - 1]
    return torch.cat([label[:, None], neg], dim=1).contiguous()


def hard_negative_loss_fwd(precision: str, q, c, selected, inv_temperature: float, sample_weight=None):
    """Returns (loss [1], row_lse [nq], row_pos [nq], scores [nq, k]) on the selected logits."""
    pc = precision_code(precision)
    dt = torch.float32 if precision == "fp32" else torch.bfloat16
    nq, d = q.shape
    K = selected.shape[1]
    dev = q.device
    scores = torch.empty((nq, K), dtype=torch.float32, device=dev)
    lse = torch.empty((nq,), dtype=torch.float32, device=dev)
    pos = torch.empty((nq,), dtype=torch.float32, device=dev)
    row_loss = torch.empty((nq,), dtype=torch.float32, device=dev)
    loss = _new_loss(dev)
    check(_lib.load().tt_hard_negative_loss_fwd(pc, _ptr(q, dt), _ptr(c, dt), _ptr(selected, torch.int64), nq, c.shape[0], K, d,
                                                inv_temperature, _ptr(sample_weight, torch.float32), _ptr(scores), _ptr(lse),
                                                _ptr(pos), _ptr(row_loss), _ptr(loss), _stream()))
    _count(2)
    return loss, lse, pos, scores


def hard_negative_loss_bwd(precision: str, q, c, selected, inv_temperature: float, scores, row_lse, sample_weight=None,
                           grad_scale: float = 1.0):
    """Returns (dq f32 [nq, d], dc f32 [nc, d])."""
    pc = precision_code(precision)
    dt = torch.float32 if precision == "fp32" else torch.bfloat16
    nq, d = q.shape
    nc = c.shape[0]
    dq = torch.empty((nq, d), dtype=torch.float32, device=q.device)
    dc = torch.empty((nc, d), dtype=torch.float32, device=q.device)
    check(_lib.load().tt_hard_negative_loss_bwd(pc, _ptr(q, dt), _ptr(c, dt), _ptr(selected, torch.int64), nq, nc, selected.shape[1], d,
                                                inv_temperature, grad_scale, _ptr(sample_weight, torch.float32),
                                                _ptr(scores, torch.float32), _ptr(row_lse, torch.float32), _ptr(dq), _ptr(dc),
                                                _stream()))
    _count(1)
    return dq, dc
