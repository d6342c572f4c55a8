"""One process per GPU: embedding tables row-sharded over the ranks with
all-to-all lookups and gradient returns, data-parallel tower MLPs (all-reduced dense
gradients), and candidate embeddings all-gathered so the in-batch negatives span the GLOBAL
batch.  The result equals the single-device TFRS loss / update on the concatenated batch.

torch.distributed (NCCL over NVLink on the box, gloo in the CPU tests) is the plumbing; every
local computation goes through ``prim`` -- the libtwotower wrappers (``ops``) in the product.
The CPU tests inject their own ``prim`` (built on the oracle) to exercise the routing logic
under gloo; this module never falls back to one by itself.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import ops as _cuda_ops
from .core import DenseGrad, GradientTape, IndexedSlices, Scalar, Tensor, Variable, config
from .layers import Dense, Layer, Sequential, _as_ids
from .models import Model


# ------------------------------------------------------------------------- collectives
class Collectives:
    """Static-shape collectives over one process group; hides backend gaps (gloo has no
    all_to_all / reduce_scatter on CPU tensors: emulated with all_gather / all_reduce)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.native = dist.get_backend(group) == "nccl"

    def all_to_all(self, x: torch.Tensor) -> torch.Tensor:
        """x [world * n, ...] -> y with y[r*n:(r+1)*n] = rank r's x[me*n:(me+1)*n]."""
        out = torch.empty_like(x)
        if self.native:
            dist.all_to_all_single(out, x, group=self.group)
            return out
        n = x.shape[0] // self.world
        bufs = [torch.empty_like(x) for _ in range(self.world)]
        dist.all_gather(bufs, x, group=self.group)
        for r in range(self.world):
            out[r * n:(r + 1) * n] = bufs[r][self.rank * n:(self.rank + 1) * n]
        return out

    def all_gather(self, x: torch.Tensor) -> torch.Tensor:
        out = torch.empty((self.world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x.contiguous(), group=self.group)
        return out

    def reduce_scatter(self, x: torch.Tensor) -> torch.Tensor:
        """x [world * n, ...] summed over ranks; this rank keeps rows [me*n, (me+1)*n)."""
        n = x.shape[0] // self.world
        if self.native:
            out = torch.empty((n,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
            dist.reduce_scatter_tensor(out, x.contiguous(), group=self.group)
            return out
        y = x.clone()
        dist.all_reduce(y, group=self.group)
        return y[self.rank * n:(self.rank + 1) * n].contiguous()

    def all_reduce_(self, x: torch.Tensor) -> torch.Tensor:
        dist.all_reduce(x, group=self.group)
        return x


# ------------------------------------------------------------------ row-sharded lookups
def bucket_capacity(batch: int, world: int, capacity_factor: Optional[float]) -> int:
    """Slots per owner in the static-shape all-to-all.  None -> batch (can never overflow)."""
    if capacity_factor is None or world == 1:
        return batch
    return min(batch, int(-(-batch * capacity_factor // world)) + 8)


def exchange_lookup(prim, coll: Collectives, table_shard: torch.Tensor, ids: torch.Tensor, capacity: int,
                    out_dtype, overflow_flag=None):
    """Rows table[ids] of a table sharded cyclically (owner = id % world, local row = id // world).
    ids [b] int64 -> (rows [b, d], ctx).  Stable partition => bit-reproducible per-owner buckets."""
    send, perm, _counts = prim.partition_ids(ids, coll.world, capacity, overflow_flag)
    recv = coll.all_to_all(send)                               # local rows wanted from me (-1 = padding)
    rows = prim.embedding_gather(table_shard, recv, out_dtype) # padding -> zero rows
    back = coll.all_to_all(rows)
    out = prim.permute_rows(back, perm, inverse=True)
    return out, (recv, perm)


def exchange_grads(prim, coll: Collectives, ctx, grad_rows: torch.Tensor, capacity: int):
    """Route d(loss)/d(rows) [b, d] back to the owners.  Returns (local_rows [world*cap] with -1
    padding, grads [world*cap, d]) == the IndexedSlices of this rank's shard."""
    recv_ids, perm = ctx
    send = prim.permute_rows(grad_rows, perm, inverse=False, out_rows=coll.world * capacity, zero_fill=True)
    recv = coll.all_to_all(send)
    return recv_ids, recv


class ShardedEmbedding(Layer):
    """tf.keras.layers.Embedding whose [input_dim, d] table is row-sharded over the group."""

    provides_tower_input = True      # Sequential may feed its output to the fused tower kernels

    def __init__(self, input_dim: int, output_dim: int, group=None, capacity_factor: Optional[float] = None,
                 name: Optional[str] = None, prim=None, seed: int = 0):
        self.coll = Collectives(group)
        self.prim = prim or _cuda_ops
        self.input_dim, self.output_dim = int(input_dim), int(output_dim)
        self.capacity_factor = capacity_factor
        self.name = name or "sharded_embedding"
        rows = -(-self.input_dim // self.coll.world)
        dev = torch.device("cuda", torch.cuda.current_device())
        g = torch.Generator(device=dev)
        g.manual_seed(config.seed * 7919 + seed * 8 + self.coll.rank)
        shard = torch.empty((rows, self.output_dim), dtype=torch.float32, device=dev)
        shard.uniform_(-0.05, 0.05, generator=g)
        self.embeddings = Variable(f"{self.name}/embeddings_shard{self.coll.rank}", shard, "table")
        self.overflow = torch.zeros(1, dtype=torch.int32, device=dev)

    @property
    def trainable_variables(self):
        return [self.embeddings]

    def load_full_table(self, table) -> None:
        """Keep rows owned by this rank (row r of the full table lives on rank r % world at r // world)."""
        t = torch.as_tensor(table, dtype=torch.float32)
        mine = t[self.coll.rank::self.coll.world]
        self.embeddings.value[:mine.shape[0]].copy_(mine.to(self.embeddings.value.device))

    def __call__(self, inputs, training: bool = False) -> Tensor:
        ids = _as_ids(inputs).reshape(-1)
        b = ids.numel()
        cap = bucket_capacity(b, self.coll.world, self.capacity_factor)
        bf16 = config.precision == "bf16"
        rows, ctx = exchange_lookup(self.prim, self.coll, self.embeddings.value, ids, cap,
                                    torch.bfloat16 if bf16 else torch.float32, self.overflow)
        if self.prim is _cuda_ops:
            GradientTape.note_sparse_lookup([(self.embeddings, ctx[0], None, "sum")])
        out = Tensor(f32=None if bf16 else rows, bf16=rows if bf16 else None, grad_formats=("f32",))

        def backward():
            if out.grad is None:
                return
            local_ids, grads = exchange_grads(self.prim, self.coll, ctx, out.grad["f32"], cap)
            self.embeddings.grad = IndexedSlices(values=local_ids, offsets=None, mode="sum", rows=grads)

        GradientTape.record(backward)
        return out

    def check_overflow(self) -> None:
        if int(self.overflow.item()) != 0:
            raise RuntimeError(f"{self.name}: an owner bucket exceeded capacity_factor={self.capacity_factor}; "
                               "rerun with capacity_factor=None")


class PeerExchange:
    """The exchanges of one row-sharded training step as kernels on a symmetric NVLink workspace
    (csrc/peer.cu, include/twotower.h) instead of NCCL collectives:

        push(candidates, ids) to every rank          -> barrier 0 -> loss forward / dQ / dC on the gathered candidates
        combine dC, each row to its owner's slot     -> barrier 1 -> the owner's tower backward folds the slots
                                                                     (was reduce-scatter)
        push gradient rows to the table-row owners,                  (was all-gather of rows)
        fold + push the dense bucket and the loss    -> barrier 2 + sum of the slots  (was 2 all-reduces)
        optimizer step                               -> barrier 3 (closes the step: peers may overwrite my
                                                                     workspace and read my tables again)

    All traffic is producer-side WRITES over NVLink (posted; SM loads from peer memory are round trips and ran
    ~4x slower here).  torch.distributed only allocates the workspace (symmetric memory rendezvous)."""

    BUCKET_FLOATS = 1 << 18          # dense-gradient bucket capacity per rank (1 MB): cfg2 needs 132 K floats

    def __init__(self, group=None):
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.ws = None
        self.tables = []              # PeerShardedEmbedding layers in lookup order of the current step
        self.step_ids = {}            # table index -> this rank's ids of the current step

    # ---- workspace
    def ensure(self, b: int, d_out: int) -> None:
        ntab = len(self.tables)
        d_emb = max([l.output_dim for l in self.tables], default=d_out)
        key = (b, d_out, d_emb, ntab)
        if self.ws is not None:
            if self.key != key:
                raise NotImplementedError(f"PeerExchange: the step shape changed {self.key} -> {key}")
            return
        import torch.distributed._symmetric_memory as symm_mem
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("PeerExchange: run one eager step before capturing a CUDA graph")
        W, bg = self.world, self.world * b
        al = lambda n: (n + 1023) // 1024 * 1024
        off = _cuda_ops.PeerWorkspace.FLAG_BYTES
        self.off_c = off; off += al(bg * d_out * 2)                    # candidates of all ranks, bf16
        self.off_ids = off; off += al(max(ntab, 1) * bg * 8)           # ids of all ranks, per table
        self.off_dc = off; off += al(bg * d_out * 4)                   # dC of MY candidates: one [b, d] slot per producing rank
        self.off_rows = off; off += al(max(ntab, 1) * bg * d_emb * 4)  # gradient rows of table rows I own, per table
        self.off_bucket = off; off += al(W * self.BUCKET_FLOATS * 4)   # dense gradients + loss: one slab per rank
        dev = torch.device("cuda", torch.cuda.current_device())
        buf = symm_mem.empty((off,), dtype=torch.uint8, device=dev)
        buf.zero_()
        torch.cuda.synchronize()
        handle = symm_mem.rendezvous(buf, self.group)
        self._handle = handle
        bases = torch.tensor([int(p) for p in handle.buffer_ptrs], dtype=torch.int64, device=dev)
        self.ws = _cuda_ops.PeerWorkspace(buf, bases, W, self.rank)
        self.key, self.b, self.d_out, self.d_emb = key, b, d_out, d_emb
        v = self.ws.view
        self.c_all = v(self.off_c, (bg, d_out), torch.bfloat16)
        self.ids_all = [v(self.off_ids + t * bg * 8, (bg,), torch.int64) for t in range(ntab)]
        self.dc_slots = v(self.off_dc, (W, b, d_out), torch.float32)
        self.rows_all = [v(self.off_rows + t * bg * d_emb * 4, (bg, d_emb), torch.float32) for t in range(ntab)]
        self.dc_mine = torch.empty((b, d_out), dtype=torch.float32, device=dev)
        self.bucket_local = torch.empty((self.BUCKET_FLOATS,), dtype=torch.float32, device=dev)
        self.bucket_sum = torch.empty((self.BUCKET_FLOATS,), dtype=torch.float32, device=dev)
        self.loss4 = torch.zeros(4, dtype=torch.float32, device=dev)
        # EXPERIMENTAL, off unless TT_DC_DIRECT=1: tensor maps over every rank's dC receive area so that the dC kernel's
        # epilogue TMA-stores straight into the owners' slots (no combine + scatter kernel).  Bit-identical to the combine
        # path on one GPU (tests/test_gpu_peer_local.py), but at N=2 one rank hung after ~2000 graph-replayed steps;
        # cause not found in round 1.
        self.dc_maps = None
        import os
        if (os.environ.get("TT_DC_DIRECT", "0") == "1" and b % 128 == 0 and d_out <= 128 and d_out % 32 == 0
                and 2 * (bg // 128) > torch.cuda.get_device_properties(dev).multi_processor_count):   # unsplit dC pass
            self.dc_maps = _cuda_ops.peer_row_maps([int(p) + self.off_dc for p in handle.buffer_ptrs], W * b, d_out, dev)
            self.dc_scratch = torch.empty((16,), dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        dist.barrier(self.group)

    # ---- step protocol
    def begin_step(self) -> None:
        self.step_ids = {}
        self.step_rows = {}

    def register_ids(self, layer, ids: torch.Tensor) -> int:
        if layer not in self.tables:
            if self.ws is not None:
                raise NotImplementedError("PeerExchange: a new sharded table appeared after the workspace was built")
            self.tables.append(layer)
        t = self.tables.index(layer)
        self.step_ids[t] = ids
        return t

    def gather_candidates(self, cm: torch.Tensor) -> torch.Tensor:
        """All ranks' candidate embeddings (and the step's ids): every rank writes its block into every copy."""
        b, d = cm.shape
        self.ensure(b, d)
        bg = self.world * b
        segs = [(cm, self.off_c + self.rank * b * d * 2)]
        for t, ids in sorted(self.step_ids.items()):
            if ids.numel() != b:
                raise NotImplementedError("PeerExchange: every sharded table must be looked up once per example")
            segs.append((ids, self.off_ids + (t * bg + self.rank * b) * 8))
        for lo in range(0, len(segs), 4):
            _cuda_ops.peer_push(self.ws, segs[lo:lo + 4])
        _cuda_ops.peer_barrier(self.ws, 0)
        notes = [(self.tables[t].embeddings, self.ids_all[t], None, "sum", (self.world, self.rank)) for t in sorted(self.step_ids)]
        if notes:
            GradientTape.note_sparse_lookup(notes)          # the optimizer's id dedup starts now, on its side stream
        return self.c_all

    def reduce_scatter_dc(self, dc_parts: torch.Tensor, as_parts: bool) -> torch.Tensor:
        """dC of my candidates: [world, b, d] stacked per-rank partials (as_parts) or their sum [b, d]."""
        _cuda_ops.peer_combine_scatter(self.ws, dc_parts, self.b, self.off_dc)
        if as_parts:
            _cuda_ops.peer_barrier(self.ws, 1)
            return self.dc_slots
        return _cuda_ops.peer_sum(self.ws, self.off_dc, self.dc_mine, slot=1, local_stride=self.b * self.d_out * 4)

    def backward_dc_direct(self, qm, c_all, inv_t, lse, label_offset, w) -> torch.Tensor:
        """dC pass + reduce-scatter in one kernel: returns the [world, b, d] slots of my candidates (after barrier 1)."""
        _cuda_ops.peer_retrieval_bwd_dc(self.ws, self.dc_maps, qm, c_all, inv_t, lse, label_offset, w, self.dc_scratch)
        _cuda_ops.peer_barrier(self.ws, 1)
        return self.dc_slots

    def table_grad(self, layer, rows: torch.Tensor) -> IndexedSlices:
        t = self.tables.index(layer)
        self.step_rows[t] = rows
        return IndexedSlices(values=self.ids_all[t], offsets=None, mode="sum", rows=self.rows_all[t],
                             shard=(self.world, self.rank))

    def finish_backward(self, dense_grads, loss: torch.Tensor):
        """dense_grads: [DenseGrad].  Returns ([summed dense gradient views], global loss [1])."""
        offs, n = _cuda_ops.bucket_layout([g.parts[0].shape for g in dense_grads])
        if n + 4 > self.BUCKET_FLOATS:
            raise NotImplementedError("PeerExchange: dense gradients exceed the bucket capacity")
        bg = self.world * self.b
        tabs = [(self.step_ids[t], self.step_rows[t], self.off_rows + t * bg * self.d_emb * 4) for t in sorted(self.step_rows)]
        if tabs:
            _cuda_ops.peer_push_rows(self.ws, tabs, self.b, self.d_emb)
        slab = self.off_bucket + self.rank * self.BUCKET_FLOATS * 4
        self.loss4[:1].copy_(loss.reshape(1))
        segs = [(self.loss4, slab + n * 4)]
        if dense_grads:
            _cuda_ops.fold_parts_into_bucket([(g.parts, g.num_parts) for g in dense_grads], out=self.bucket_local)
            segs.insert(0, (self.bucket_local[:n], slab))
        _cuda_ops.peer_push(self.ws, segs)
        _cuda_ops.peer_sum(self.ws, self.off_bucket, self.bucket_sum[:n + 4], slot=2, local_stride=self.BUCKET_FLOATS * 4)
        views = [self.bucket_sum[o:o + g.parts[0].numel()].view(g.parts[0].shape) for o, g in zip(offs, dense_grads)]
        return views, self.bucket_sum[n:n + 1]

    def end_step(self) -> None:
        _cuda_ops.peer_barrier(self.ws, 3)


class PeerShardedEmbedding(Layer):
    """tf.keras.layers.Embedding whose [input_dim, d] table is row-sharded over the group (owner = id % world)
    in SYMMETRIC memory: every rank maps its peers' shards (torch.distributed._symmetric_memory, NVLink P2P),
    so a lookup is a plain gather -- the tower kernel loads each row straight from its owner over NVLink, fused
    with the pooling and the tower MLP; there is no id partition, no all-to-all and no row permutation.
    Backward: the (id, gradient row) pairs of all ranks are all-gathered and each owner's optimizer launch
    keeps the entries it owns (tt_sparse_var.shard).

    Ordering across ranks: a step's remote reads of a shard happen before its owner's optimizer launch of
    that step only because collectives sit in between on every rank's stream (candidate all-gather, ...),
    and the owner's update is visible to the next step's remote reads because DataParallelModel.train_step
    ends with an all-reduce after the optimizer launch."""

    combiner = None

    def __init__(self, input_dim: int, output_dim: int, group=None, name: Optional[str] = None, seed: int = 0):
        import torch.distributed._symmetric_memory as symm_mem
        self.coll = Collectives(group)
        self.input_dim, self.output_dim = int(input_dim), int(output_dim)
        self.name = name or "peer_sharded_embedding"
        world, rank = self.coll.world, self.coll.rank
        rows = -(-self.input_dim // world)
        dev = torch.device("cuda", torch.cuda.current_device())
        shard = symm_mem.empty((rows, self.output_dim), dtype=torch.float32, device=dev)
        g = torch.Generator(device=dev)
        g.manual_seed(config.seed * 7919 + seed * 8 + rank)
        shard.uniform_(-0.05, 0.05, generator=g)
        handle = symm_mem.rendezvous(shard, group if group is not None else dist.group.WORLD)
        self._handle = handle
        ptrs = torch.tensor([int(p) for p in handle.buffer_ptrs], dtype=torch.int64, device=dev)
        self.table = _cuda_ops.PeerTable(ptrs, shard, self.input_dim, world, rank)
        self.embeddings = Variable(f"{self.name}/embeddings_shard{rank}", shard, "table")
        self._ids_all = None
        self.exchange = None          # PeerExchange: ids / gradient rows travel through the symmetric workspace

    @property
    def trainable_variables(self):
        return [self.embeddings]

    def load_full_table(self, table) -> None:
        t = torch.as_tensor(table, dtype=torch.float32)
        mine = t[self.coll.rank::self.coll.world]
        self.embeddings.value[:mine.shape[0]].copy_(mine.to(self.embeddings.value.device))
        torch.cuda.synchronize()
        dist.barrier(self.coll.group)

    def _feature(self, inputs):
        return (self.table, _as_ids(inputs).reshape(-1), None, "sum")

    def _tower_features(self, inputs):
        return [self], [self._feature(inputs)]

    def _lookup_note(self, feat):
        if self.exchange is not None:
            self.exchange.register_ids(self, feat[1])              # announced after the id exchange (gather_candidates)
            return None
        self._ids_all = self.coll.all_gather(feat[1])              # global ids of every rank's batch
        return (self.embeddings, self._ids_all, None, "sum", (self.coll.world, self.coll.rank))

    def _make_grad(self, feat, rows) -> IndexedSlices:
        if self.exchange is not None:
            return self.exchange.table_grad(self, rows)
        ids_all = self._ids_all if self._ids_all is not None else self.coll.all_gather(feat[1])
        self._ids_all = None
        return IndexedSlices(values=ids_all, offsets=None, mode="sum", rows=self.coll.all_gather(rows),
                             shard=(self.coll.world, self.coll.rank))

    def __call__(self, inputs, training: bool = False) -> Tensor:
        from .layers import _tower_input
        return _tower_input([self], [inputs])


# ------------------------------------------------------------------- global negatives
def global_retrieval(task, q: Tensor, c: Tensor, inv_t: float, w, logq, ids, prim=None) -> Scalar:
    """tfrs.tasks.Retrieval with the candidates all-gathered over task.process_group: rank r's
    queries score against all world*b candidates, positives at columns [r*b, (r+1)*b).
    Loss reported = this rank's share; sum over ranks == single-device loss on the global batch."""
    prim = prim or _cuda_ops
    coll = Collectives(task.process_group)
    prec = config.precision
    qm = q.f32 if prec == "fp32" else q.bf16
    cm = c.f32 if prec == "fp32" else c.bf16
    nq = qm.shape[0]
    ex = getattr(task, "exchange", None)
    if ex is not None and not (prec == "bf16" and prim is _cuda_ops):
        raise NotImplementedError("the peer-memory exchange runs the bf16 kernels only")
    c_all = ex.gather_candidates(cm) if ex is not None else coll.all_gather(cm)      # [world*b, d]
    label_offset = coll.rank * nq
    logq_all = None if logq is None else coll.all_gather(logq)
    ids_all = None if ids is None else coll.all_gather(ids)
    fused = (prec == "bf16" and prim is _cuda_ops and GradientTape.current() is not None and logq_all is None
             and ids_all is None and ("parts" in q.grad_formats or "parts" in c.grad_formats)
             and prim.retrieval_fwd_dq_supported(nq, c_all.shape[0], qm.shape[1]))
    if fused:            # the forward pass also accumulates dQ: the backward is the dC pass only
        loss, lse, _pos, dq_fused, _fwd_ws = prim.retrieval_loss_fwd_dq(qm, c_all, inv_t, label_offset, w, fork=True)
    else:
        loss, lse, _pos = prim.retrieval_loss_fwd(prec, qm, c_all, inv_t, label_offset, w, logq_all, ids_all)

    def backward():
        bf = prec == "bf16"
        if bf and prim is _cuda_ops and ("parts" in q.grad_formats or "parts" in c.grad_formats):
            if fused:
                dq_parts = dq_fused.reshape(1, *dq_fused.shape)
                if ex is not None and ex.dc_maps is not None and "parts" in c.grad_formats:
                    # the dC kernel's epilogue writes every row block straight into its owner's slot over NVLink
                    c.grad = dict(parts=ex.backward_dc_direct(qm, c_all, inv_t, lse, label_offset, w))
                    prim.join_side_work()
                    if "parts" in q.grad_formats:
                        q.grad = dict(parts=dq_parts)
                    else:
                        f, b = prim.combine_parts(dq_parts, True, "bf16" in q.grad_formats)
                        q.grad = dict(f32=f, bf16=b)
                    return
                _none, dc_parts = prim.retrieval_loss_bwd_parts(qm, c_all, inv_t, lse, label_offset, w, None, None, 1.0,
                                                                want_dq=False)
                prim.join_side_work()                # the forked loss summation joins AFTER the dC launch
            else:
                dq_parts, dc_parts = prim.retrieval_loss_bwd_parts(qm, c_all, inv_t, lse, label_offset, w, logq_all, ids_all, 1.0)
            if "parts" in q.grad_formats:
                q.grad = dict(parts=dq_parts)
            else:
                f, b = prim.combine_parts(dq_parts, True, "bf16" in q.grad_formats)
                q.grad = dict(f32=f, bf16=b)
            if ex is not None:
                dc = ex.reduce_scatter_dc(dc_parts, as_parts="parts" in c.grad_formats)   # [world, b, d] slots if parts
            else:
                dc_full, _ = prim.combine_parts(dc_parts, True, False)
                dc = coll.reduce_scatter(dc_full)                 # every rank's partial for my candidates
            if "parts" in c.grad_formats:
                c.grad = dict(parts=dc if dc.dim() == 3 else dc.reshape(1, *dc.shape))
            else:
                c.grad = dict(f32=dc, bf16=prim.cast_f32_to_bf16(dc) if "bf16" in c.grad_formats else None)
            return
        if ex is not None:
            raise NotImplementedError("the peer-memory exchange needs the fused tower kernels (split-partial gradients)")
        r = prim.retrieval_loss_bwd(prec, qm, c_all, inv_t, lse, label_offset, w, logq_all, ids_all, 1.0,
                                    want_bf16=(bf and "bf16" in q.grad_formats, False))
        q.grad = dict(f32=r["dq"], bf16=r["dq_bf16"])
        dc = coll.reduce_scatter(r["dc"])                         # every rank's partial for my candidates
        dc_b = prim.cast_f32_to_bf16(dc) if (bf and "bf16" in c.grad_formats) else None
        c.grad = dict(f32=dc, bf16=dc_b)

    GradientTape.record(backward)
    return Scalar(loss)


class DataParallelModel(Model):
    """tfrs.models.Model whose Dense gradients are summed over the group before the update (the
    tables are sharded, their gradients were already routed to the owners)."""

    def __init__(self, group=None, name: Optional[str] = None):
        super().__init__(name)
        self.process_group = group
        self._coll = Collectives(group)
        self.exchange = None          # PeerExchange (build_sharded_two_tower(peer="exchange"))

    # every rank holds different table shards: one checkpoint file per rank
    def _rank_path(self, path) -> str:
        return f"{path}.rank{self._coll.rank}of{self._coll.world}.npz"

    def save_weights(self, path) -> None:
        super().save_weights(self._rank_path(path))

    def load_weights(self, path) -> None:
        super().load_weights(self._rank_path(path))

    def train_step(self, inputs):
        if self.optimizer is None:
            raise RuntimeError("call model.compile(optimizer=...) before train_step")
        self.optimizer.begin_step()
        ex = self.exchange
        if ex is not None:
            ex.begin_step()
        with GradientTape() as tape:
            tape.on_sparse_lookup = self.optimizer.prepare_sparse
            loss = self.compute_loss(inputs, training=True)
            self.optimizer.join_prepare()
            variables = self.trainable_variables
            grads = tape.gradient(loss, variables)
        dense = [(i, g) for i, g in enumerate(grads) if isinstance(g, DenseGrad)]
        if ex is not None:
            views, total = ex.finish_backward([g for _, g in dense], loss.value)
            for (i, g), v in zip(dense, views):
                grads[i] = DenseGrad(v.reshape((1,) + tuple(v.shape)), 1)
            self.optimizer.apply_gradients(zip(grads, variables))
            ex.end_step()
            return {"loss": total, "local_loss": loss.value, "regularization_loss": torch.zeros_like(total), "total_loss": total}
        if dense:
            # every dense gradient folded into ONE flat bucket (one launch) and summed over the ranks by a
            # single latency-bound all-reduce per step
            if any(g.num_parts > 1 for _, g in dense):
                bucket, views = _cuda_ops.fold_parts_into_bucket([(g.parts, g.num_parts) for _, g in dense])
            else:
                bucket = torch.cat([g.parts[0].reshape(-1) for _, g in dense])
                views, off = [], 0
                for _, g in dense:
                    n = g.parts[0].numel()
                    views.append(bucket[off:off + n].view(g.parts[0].shape))
                    off += n
            self._coll.all_reduce_(bucket)
            for (i, g), v in zip(dense, views):
                grads[i] = DenseGrad(v.reshape((1,) + tuple(v.shape)), 1)
        self.optimizer.apply_gradients(zip(grads, variables))
        total = loss.value.clone()
        self._coll.all_reduce_(total)
        return {"loss": total, "local_loss": loss.value, "regularization_loss": torch.zeros_like(total), "total_loss": total}


def build_sharded_two_tower(cfg, group, lr: float = 0.001, capacity_factor: Optional[float] = 2.0, peer=True):
    """The bench model at N > 1: ID-only two-tower of `cfg`, tables row-sharded, global negatives.
    peer="exchange": shards in symmetric memory, lookups are P2P gathers fused into the tower kernel AND every
    exchange of the step (candidates, dC, gradient rows, dense gradients, loss) is a peer-memory kernel
    (PeerExchange; bf16 only); peer=True: P2P lookups, NCCL collectives for the rest; peer=False: NCCL
    all-to-all lookups (ShardedEmbedding)."""
    from . import optimizers, tasks

    class ShardedTwoTower(DataParallelModel):
        def __init__(self):
            super().__init__(group)
            def tower(vocab, seed):
                layers = [PeerShardedEmbedding(vocab, cfg.dim, group, seed=seed) if peer else
                          ShardedEmbedding(vocab, cfg.dim, group, capacity_factor, seed=seed)]
                for j, u in enumerate(cfg.mlp):
                    layers.append(Dense(u, "relu" if j < len(cfg.mlp) - 1 else None))
                return Sequential(layers)
            self.user_model = tower(cfg.v_user, 1)
            self.item_model = tower(cfg.v_item, 2)
            self.task = tasks.Retrieval(temperature=cfg.temperature, process_group=group)
            if peer == "exchange":
                self.exchange = PeerExchange(group)
                self.task.exchange = self.exchange
                self.user_model.layers[0].exchange = self.exchange
                self.item_model.layers[0].exchange = self.exchange

        def compute_loss(self, features, training=False):
            return self.task(self.user_model(features["user_id_encoded"]), self.item_model(features["item_id_encoded"]))

    model = ShardedTwoTower()
    model.compile(optimizer=optimizers.Adagrad(learning_rate=lr))
    return model
