"""Minimal tape, tensor and variable types behind the TFRS-shaped Python surface.

A ``Tensor`` is a handle to device buffers produced by libtwotower kernels (torch tensors are
only the memory owners).  ``GradientTape`` records one backward closure per layer/task call,
mirroring what tfrs.models.Model.train_step does with tf.GradientTape.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional

import numpy as np
import torch


class _Config:
    precision = "bf16"          # "fp32" (CUDA-core, 1e-5 parity) | "bf16" (tcgen05, 2e-2 parity)
    seed = 0


config = _Config()


def set_precision(precision: str) -> None:
    if precision not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    config.precision = precision


def set_seed(seed: int) -> None:
    config.seed = int(seed)


def device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("two_tower_b200 needs a CUDA device (sm_100a); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


class Tensor:
    """Activations of one layer: an fp32 copy and/or a bf16 copy.
    ``grad_formats`` tells the consumer which gradient representations the producer needs
    ("f32", "bf16", or "parts" = stacked fp32 split partials the producer folds itself).
    A tensor may be *pending*: its buffers are produced on first access, together with every
    other pending tower output, by one fused launch (layers.PendingTowers)."""

    def __init__(self, f32=None, bf16=None, grad_formats=("f32",), producer=None, pending=None, shape=None):
        self._f32: Optional[torch.Tensor] = f32
        self._bf16: Optional[torch.Tensor] = bf16
        self._pending = pending
        self._shape = shape
        self.grad_formats = tuple(grad_formats)
        self.producer = producer
        self.grad = None          # dict(f32=..., bf16=..., parts=...) set by the consumer's backward
        self.relu_output = False  # True when this is the output of a relu Dense
        self.producer_needs_grad = True   # False for constants fed in by the user

    def _force(self):
        if self._pending is not None:
            self._pending.flush()

    @property
    def f32(self) -> Optional[torch.Tensor]:
        self._force()
        return self._f32

    @f32.setter
    def f32(self, v):
        self._f32 = v

    @property
    def bf16(self) -> Optional[torch.Tensor]:
        self._force()
        return self._bf16

    @bf16.setter
    def bf16(self, v):
        self._bf16 = v

    @property
    def shape(self):
        if self._pending is not None:
            return tuple(self._shape)
        t = self._f32 if self._f32 is not None else self._bf16
        return tuple(t.shape)

    def torch(self) -> torch.Tensor:
        """fp32 view of the activations (device)."""
        return self.f32 if self.f32 is not None else self.bf16.float()

    def numpy(self) -> np.ndarray:
        return self.torch().detach().cpu().numpy()


class Scalar:
    """A device scalar (loss) with a tape node."""

    def __init__(self, value: torch.Tensor):
        self.value = value        # [1] fp32 device tensor

    def item(self) -> float:
        return float(self.value.item())

    def numpy(self):
        return self.value.detach().cpu().numpy()[0]


@dataclass
class IndexedSlices:
    """Sparse gradient of an embedding table: un-expanded rows + CSR membership
    (tf.IndexedSlices with duplicates)."""
    values: torch.Tensor              # ids [nnz] int64
    offsets: Optional[torch.Tensor]   # CSR offsets [rows+1] or None (one id per row)
    mode: str                         # "sum" | "mean"
    rows: torch.Tensor                # [num_rows, d] fp32 upstream gradient
    shard: Optional[tuple] = None     # (world, rank): values are GLOBAL ids of a row-sharded table, the variable
                                      # is this rank's shard (entries of other owners are skipped)


@dataclass
class DenseGrad:
    parts: torch.Tensor               # [P, rows, cols] fp32 partial sums
    num_parts: int


class Variable:
    def __init__(self, name: str, value: torch.Tensor, kind: str, l2: float = 0.0):
        self.name = name
        self.value = value            # fp32 master copy
        self.kind = kind              # "table" | "kernel" | "bias"
        self.l2 = float(l2)
        self.want_shadows = False     # keep a bf16 copy of the value in step with it
        self.shadow = None            # bf16 [rows, cols]   (kernel only, bf16 precision)
        self.grad = None
        self.slots = {}               # optimizer state

    @property
    def shape(self):
        return tuple(self.value.shape)

    def numpy(self):
        return self.value.detach().cpu().numpy()

    def assign(self, array) -> None:
        from . import ops
        a = torch.as_tensor(np.asarray(array), dtype=torch.float32)
        if tuple(a.shape) != self.shape:
            raise ValueError(f"assign: shape {tuple(a.shape)} != {self.shape}")
        self.value.copy_(a.to(self.value.device))
        self.refresh_shadows()

    def refresh_shadows(self) -> None:
        """bf16 copy of the value, rewritten IN PLACE once it exists: captured CUDA graphs and the optimizer kernel hold
        its address."""
        from . import ops
        if self.want_shadows:
            self.shadow = ops.cast_f32_to_bf16(self.value, out=self.shadow)

    def assign_slot(self, key: str, array) -> None:
        """Set an optimizer slot, in place when it already exists with the same shape (same reason)."""
        t = torch.as_tensor(np.asarray(array)) if not isinstance(array, torch.Tensor) else array
        cur = self.slots.get(key)
        if isinstance(cur, torch.Tensor) and tuple(cur.shape) == tuple(t.shape):
            cur.copy_(t.to(device=cur.device, dtype=cur.dtype))
        else:
            self.slots[key] = t.to(self.value.device).clone()


class GradientTape:
    _stack: List["GradientTape"] = []

    def __init__(self):
        self.nodes: List[Callable[[], None]] = []
        # called as on_sparse_lookup([(variable, values, offsets, mode)]) whenever embedding tables are read
        # under this tape: lets the optimizer start the id dedup while the forward pass is still running
        self.on_sparse_lookup: Optional[Callable] = None

    def __enter__(self):
        GradientTape._stack.append(self)
        return self

    def __exit__(self, *exc):
        GradientTape._stack.pop()
        from . import ops
        ops.join_side_work()          # a forward whose backward never ran may have left a forked fold in flight
        return False

    @staticmethod
    def current() -> Optional["GradientTape"]:
        return GradientTape._stack[-1] if GradientTape._stack else None

    @staticmethod
    def note_sparse_lookup(lookups, fused: bool = False):
        """fused: the caller's own kernel performs the id dedup (the fused tower forward); the callback then only
        registers the lookups and returns one sparse-optimizer workspace per lookup (None without a listener)."""
        t = GradientTape.current()
        if t is not None and t.on_sparse_lookup is not None:
            return t.on_sparse_lookup(lookups, fused=True) if fused else t.on_sparse_lookup(lookups)
        return None

    @staticmethod
    def record(fn: Callable[[], None]) -> None:
        t = GradientTape.current()
        if t is not None:
            t.nodes.append(fn)

    def gradient(self, target, variables):
        """Run the recorded backward closures in reverse order; returns [v.grad for v in variables]."""
        for v in variables:
            v.grad = None
        for fn in reversed(self.nodes):
            fn()
        self.nodes.clear()
        return [v.grad for v in variables]
