"""Keras-2.15-semantics optimizers applied by libtwotower kernels: sparse
(dedup + row-wise update) for embedding tables, dense for Dense kernels/biases."""
from __future__ import annotations

import math
from typing import Iterable, Tuple

import torch

from . import ops
from .core import DenseGrad, IndexedSlices, Variable


class Optimizer:
    def __init__(self):
        self.iterations = 0

    def apply_gradients(self, grads_and_vars: Iterable[Tuple[object, Variable]]) -> None:
        self.iterations += 1
        for g, v in grads_and_vars:
            if g is None:
                continue
            if isinstance(g, IndexedSlices):
                self._apply_sparse(g, v)
            elif isinstance(g, DenseGrad):
                self._apply_dense(g, v)
            else:
                raise TypeError(f"unsupported gradient type {type(g)} for {v.name}")

    @staticmethod
    def _sparse_ws(var: Variable, nnz: int) -> ops.SparseWorkspace:
        ws = var.slots.get("_sparse_ws")
        d = var.value.shape[1]
        if ws is None or not ws.fits(nnz, d):
            ws = ops.SparseWorkspace(max(nnz, 1), d, var.value.device)
            var.slots["_sparse_ws"] = ws
        return ws


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

    def _apply_sparse(self, g: IndexedSlices, var: Variable) -> None:
        ws = self._sparse_ws(var, g.values.numel())
        ops.sparse_adagrad_update(var.value, self._acc(var), g.values, g.offsets, g.mode, g.rows,
                                  self.learning_rate, self.epsilon, ws, var.slots.get("_first_flag"))

    def _apply_dense(self, g: DenseGrad, var: Variable) -> None:
        ops.dense_adagrad_update(var.value, self._acc(var), g.parts, g.num_parts, self.learning_rate,
                                 self.epsilon, var.l2, var.shadow if var.want_shadows else None)


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

    def _mv(self, var: Variable):
        if "m" not in var.slots:
            var.slots["m"] = torch.zeros_like(var.value)
            var.slots["v"] = torch.zeros_like(var.value)
        return var.slots["m"], var.slots["v"]

    def _apply_dense(self, g: DenseGrad, var: Variable) -> None:
        m, v = self._mv(var)
        ops.dense_adam_update(var.value, m, v, g.parts, g.num_parts, self._alpha(), self.beta_1, self.beta_2,
                              self.epsilon, var.l2, var.shadow if var.want_shadows else None)

    def _apply_sparse(self, g: IndexedSlices, var: Variable) -> None:
        if not self.lazy:
            raise NotImplementedError(
                "Keras Adam on an embedding table is a dense whole-table update (SURVEY.md A.6); "
                "use LazyAdam (touched rows only) or Adagrad for tables")
        m, v = self._mv(var)
        ws = self._sparse_ws(var, g.values.numel())
        ops.sparse_lazy_adam_update(var.value, m, v, g.values, g.offsets, g.mode, g.rows, self._alpha(),
                                    self.beta_1, self.beta_2, self.epsilon, ws, var.slots.get("_first_flag"))


class LazyAdam(Adam):
    """tfa.optimizers.LazyAdam semantics: Adam arithmetic on the touched rows only."""
    lazy = True
