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
        self._ids_all = self.coll.all_gather(feat[1])              # global ids of every rank's batch
        return (self.embeddings, self._ids_all, None, "sum", (self.coll.world, self.coll.rank))

    def _make_grad(self, feat, rows) -> IndexedSlices:
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
    c_all = coll.all_gather(cm)                                   # [world*b, d]
    label_offset = coll.rank * nq
    logq_all = None if logq is None else coll.all_gather(logq)
    ids_all = None if ids is None else coll.all_gather(ids)
    loss, lse, _pos = prim.retrieval_loss_fwd(prec, qm, c_all, inv_t, label_offset, w, logq_all, ids_all)

    def backward():
        bf = prec == "bf16"
        if bf and prim is _cuda_ops and ("parts" in q.grad_formats or "parts" in c.grad_formats):
            dq_parts, dc_parts = prim.retrieval_loss_bwd_parts(qm, c_all, inv_t, lse, label_offset, w, logq_all, ids_all, 1.0)
            if "parts" in q.grad_formats:
                q.grad = dict(parts=dq_parts)
            else:
                f, b = prim.combine_parts(dq_parts, True, "bf16" in q.grad_formats)
                q.grad = dict(f32=f, bf16=b)
            dc_full, _ = prim.combine_parts(dc_parts, True, False)
            dc = coll.reduce_scatter(dc_full)                     # every rank's partial for my candidates
            if "parts" in c.grad_formats:
                c.grad = dict(parts=dc.reshape(1, *dc.shape))
            else:
                c.grad = dict(f32=dc, bf16=prim.cast_f32_to_bf16(dc) if "bf16" in c.grad_formats else None)
            return
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

    def train_step(self, inputs):
        if self.optimizer is None:
            raise RuntimeError("call model.compile(optimizer=...) before train_step")
        self.optimizer.begin_step()
        with GradientTape() as tape:
            tape.on_sparse_lookup = self.optimizer.prepare_sparse
            loss = self.compute_loss(inputs, training=True)
            variables = self.trainable_variables
            grads = tape.gradient(loss, variables)
        dense = [(i, g) for i, g in enumerate(grads) if isinstance(g, DenseGrad)]
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


def build_sharded_two_tower(cfg, group, lr: float = 0.001, capacity_factor: Optional[float] = 2.0, peer: bool = True):
    """The bench model at N > 1: ID-only two-tower of `cfg`, tables row-sharded, global negatives.
    peer=True: shards in symmetric memory, lookups are P2P gathers fused into the tower kernel
    (PeerShardedEmbedding); peer=False: NCCL all-to-all lookups (ShardedEmbedding)."""
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

        def compute_loss(self, features, training=False):
            return self.task(self.user_model(features["user_id_encoded"]), self.item_model(features["item_id_encoded"]))

    model = ShardedTwoTower()
    model.compile(optimizer=optimizers.Adagrad(learning_rate=lr))
    return model
