"""Keras-2.15-semantics optimizers applied by libtwotower kernels: sparse
(dedup + row-wise update) for embedding tables, dense for Dense kernels/biases."""
from __future__ import annotations

import os

import math
from typing import Iterable, Tuple

import torch

from . import ops
from .core import DenseGrad, IndexedSlices, Variable


_INLINE_PREPARE = os.environ.get("TT_PREPARE_INLINE", "0") == "1"      # measured slower (below): kept for A/B runs


class Optimizer:
    """apply_gradients runs the whole step as ONE launch (all dense variables + all tables) after a
    hash-insert launch that `prepare_sparse` enqueues on a side stream at lookup time (it needs only
    the ids), so it overlaps the forward pass (tt_optimizer_prepare_sparse / tt_*_step)."""

    kind = "adagrad"

    def __init__(self):
        self.iterations = 0        # host mirror of the step count (a GraphedStep adds 1 per replay)
        self._prepared = {}        # id(variable) -> values tensor whose insert is in flight on the side stream
        self._prepared_vars = []
        self._side = None
        self._side_busy = False

    # ---- overridden by the concrete optimizers
    def _dense_item(self, g: DenseGrad, v: Variable):
        raise NotImplementedError

    def _sparse_item(self, v: Variable, values, offsets, mode, rows):
        raise NotImplementedError

    def _hyper(self):
        raise NotImplementedError

    def _check_sparse_supported(self) -> None:
        pass

    # ---- the step
    def begin_step(self) -> None:
        """Called by Model.train_step before the forward.  If an earlier step announced lookups and never
        applied them (compute_loss raised), their hash tables still hold the ids: rebuild those workspaces."""
        if self._prepared:
            if self._side_busy:
                torch.cuda.current_stream().wait_stream(self._side)
                self._side_busy = False
            for ws_owner in self._prepared_vars:
                ws_owner.slots.pop("_sparse_ws", None)
            self._prepared.clear()
        self._prepared_vars = []

    def prepare_sparse(self, lookups, fused: bool = False):
        """lookups: [(variable, values, offsets, mode)].  Enqueue the id dedup (hash insert) of these tables
        on a side stream; apply_gradients joins it.  fused=True: nothing is launched -- the caller's kernel (the
        fused tower forward) does the insert -- and the workspace of every lookup is returned."""
        self._check_sparse_supported()
        if fused:
            out = []
            for look in lookups:
                var, values, offsets, mode = look[:4]
                if len(look) > 4 and look[4] is not None or offsets is not None or values.numel() == 0:
                    raise ValueError("fused id dedup: unsharded ID lookups only")
                if id(var) in self._prepared:
                    raise NotImplementedError("an embedding table used twice in one step is not supported")
                out.append(self._sparse_ws(var, values.numel()))
                self._prepared[id(var)] = values
                self._prepared_vars.append(var)
            return out
        items = []
        for look in lookups:
            var, values, offsets, mode = look[:4]
            shard = look[4] if len(look) > 4 else None
            if id(var) in self._prepared:
                raise NotImplementedError("an embedding table used twice in one step is not supported")
            if values.numel() == 0:
                continue
            items.append(self._sparse_item(var, values, offsets, mode, None) + (shard,))
            self._prepared[id(var)] = values
            self._prepared_vars.append(var)
        if not items:
            return
        if _INLINE_PREPARE:
            # On the step's own stream, ahead of the tower kernels: no side-stream join later (a join turns the launch
            # behind it into a full dependency: 6 us of idle before the dC pass against ~1 us for a PDL edge), but the
            # hash inserts take ~12 us in series.  Measured cfg2: 163.9 us/step inline against 155.2 on the side stream.
            ops.sparse_prepare(items)
            return
        cur = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            ops.sparse_prepare(items)
        self._side_busy = True

    def join_prepare(self) -> None:
        """Make the current stream wait for the side-stream id dedup now.  Model.train_step calls this between the
        forward and the backward pass: the dedup has long finished there, and the optimizer launch then follows the
        backward tower kernel directly (a programmatic dependent launch) instead of an event wait."""
        if self._side_busy:
            torch.cuda.current_stream().wait_stream(self._side)
            self._side_busy = False

    def apply_gradients(self, grads_and_vars: Iterable[Tuple[object, Variable]]) -> None:
        self.iterations += 1
        dense, sparse, late = [], [], []
        for g, v in grads_and_vars:
            if g is None:
                continue
            if isinstance(g, IndexedSlices):
                if g.values.numel() == 0:
                    continue
                item = self._sparse_item(v, g.values, g.offsets, g.mode, g.rows) + (g.shard,)
                if self._prepared.get(id(v)) is not g.values:
                    late.append(item)          # not announced at lookup time: dedup now, on this stream
                sparse.append(item)
            elif isinstance(g, DenseGrad):
                dense.append(self._dense_item(g, v))
            else:
                raise TypeError(f"unsupported gradient type {type(g)} for {v.name}")
        if sparse:
            self._check_sparse_supported()
        if self._side_busy:
            torch.cuda.current_stream().wait_stream(self._side)
            self._side_busy = False
        self._prepared.clear()
        self._prepared_vars = []
        if late:
            ops.sparse_prepare(late)
        hyper = self._hyper()
        for lo in range(0, max(len(dense), 1), 16):
            chunk_d = dense[lo:lo + 16]
            chunk_s = sparse if lo == 0 else []
            if chunk_d or chunk_s:
                ops.optimizer_step(self.kind, chunk_d, chunk_s[:8], hyper)
        for lo in range(8, len(sparse), 8):
            ops.optimizer_step(self.kind, [], sparse[lo:lo + 8], hyper)

    @staticmethod
    def _sparse_ws(var: Variable, nnz: int) -> ops.SparseWorkspace:
        ws = var.slots.get("_sparse_ws")
        d = var.value.shape[1]
        if ws is None or not ws.fits(nnz, d):
            ws = ops.SparseWorkspace(max(nnz, 1), d, var.value.device)
            var.slots["_sparse_ws"] = ws
        return ws


    def set_iterations(self, t: int) -> None:
        self.iterations = int(t)


class Adagrad(Optimizer):
    """tf.keras.optimizers.Adagrad(learning_rate=0.001, initial_accumulator_value=0.1,
    epsilon=1e-7): acc += g^2; w -= lr * g / sqrt(acc + eps)  (learning_rate default from
    /root/reference/configs/data_config.yaml:63)."""

    def __init__(self, learning_rate: float = 0.001, initial_accumulator_value: float = 0.1,
                 epsilon: float = 1e-7):
        super().__init__()
        self.learning_rate = float(learning_rate)
        self.initial_accumulator_value = float(initial_accumulator_value)
        self.epsilon = float(epsilon)

    def _acc(self, var: Variable) -> torch.Tensor:
        a = var.slots.get("accumulator")
        if a is None:
            a = torch.full_like(var.value, self.initial_accumulator_value)
            var.slots["accumulator"] = a
        return a

    def _dense_item(self, g: DenseGrad, v: Variable):
        return (v.value, self._acc(v), None, g.parts, g.num_parts, v.l2, v.shadow if v.want_shadows else None)

    def _sparse_item(self, v: Variable, values, offsets, mode, rows):
        return (v.value, self._acc(v), None, values, offsets, mode, rows, self._sparse_ws(v, values.numel()),
                v.slots.get("_first_flag"))

    def _hyper(self):
        return (self.learning_rate, self.epsilon)


class Adam(Optimizer):
    """tf.keras.optimizers.Adam for dense variables.  Keras' sparse Adam rewrites the WHOLE table
    every step; the row-wise variant north_star names is ``LazyAdam``."""

    lazy = False

    def __init__(self, learning_rate: float = 0.001, beta_1: float = 0.9, beta_2: float = 0.999,
                 epsilon: float = 1e-7):
        super().__init__()
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = map(float, (learning_rate, beta_1, beta_2, epsilon))

    def _alpha(self) -> float:
        t = self.iterations
        return self.learning_rate * math.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t)

    # The step count and the bias-corrected step size live on the DEVICE: tt_adam_bias_correction advances the
    # counter and writes alpha_t at the head of every step, inside the captured graph as well, so that a replayed
    # step uses the alpha of ITS iteration (a host scalar would freeze alpha at capture time).
    def _device_state(self, device):
        st = getattr(self, "_dev_state", None)
        if st is None or st[0].device != device:
            st = (torch.tensor([self.iterations], dtype=torch.int64, device=device),
                  torch.zeros(1, dtype=torch.float32, device=device))
            self._dev_state = st
        return st

    def set_iterations(self, t: int) -> None:
        """Restore the step count (checkpoint load): host mirror and device counter, in place."""
        self.iterations = int(t)
        st = getattr(self, "_dev_state", None)
        if st is not None:
            st[0].fill_(int(t))

    def apply_gradients(self, grads_and_vars) -> None:
        gv = list(grads_and_vars)
        dev = next((v.value.device for _g, v in gv), None)
        if dev is not None:
            step_dev, alpha_dev = self._device_state(dev)
            ops.adam_bias_correction(step_dev, alpha_dev, self.learning_rate, self.beta_1, self.beta_2)
        super().apply_gradients(gv)

    def _mv(self, var: Variable):
        if "m" not in var.slots:
            var.slots["m"] = torch.zeros_like(var.value)
            var.slots["v"] = torch.zeros_like(var.value)
        return var.slots["m"], var.slots["v"]

    kind = "lazy_adam"      # Adam arithmetic; tables only with lazy = True

    def _dense_item(self, g: DenseGrad, v: Variable):
        return (v.value, *self._mv(v), g.parts, g.num_parts, v.l2, v.shadow if v.want_shadows else None)

    def _sparse_item(self, v: Variable, values, offsets, mode, rows):
        return (v.value, *self._mv(v), values, offsets, mode, rows, self._sparse_ws(v, values.numel()),
                v.slots.get("_first_flag"))

    def _hyper(self):
        st = getattr(self, "_dev_state", None)
        return (st[1] if st is not None else self._alpha(), self.beta_1, self.beta_2, self.epsilon)

    def _check_sparse_supported(self) -> None:
        if not self.lazy:
            raise NotImplementedError(
                "Keras Adam on an embedding table is a dense whole-table update (SURVEY.md A.6); "
                "use LazyAdam (touched rows only) or Adagrad for tables")


class LazyAdam(Adam):
    """tfa.optimizers.LazyAdam semantics: Adam arithmetic on the touched rows only."""
    lazy = True
