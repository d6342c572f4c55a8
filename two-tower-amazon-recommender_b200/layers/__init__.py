"""Tower layers: Embedding, pooled multi-hot Embedding, Dense, Sequential -- the
tf.keras.layers surface a TFRS two-tower model is built from (sizes from
/root/reference/configs/data_config.yaml:55-59, ids from
/root/reference/src/data/preprocessor.py:481-489)."""
from __future__ import annotations

import math
import os
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from .. import ops
from ..core import DenseGrad, GradientTape, IndexedSlices, Tensor, Variable, config, device

_layer_counter = [0]


def _next_seed() -> int:
    _layer_counter[0] += 1
    return config.seed * 1000003 + _layer_counter[0]


def _as_ids(x) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.as_tensor(np.asarray(x))
    if t.dtype != torch.int64:
        t = t.to(torch.int64)
    if not t.is_cuda:
        t = t.to(device(), non_blocking=True)
    return t.contiguous()


class Layer:
    name = "layer"

    @property
    def trainable_variables(self):
        return []

    @property
    def losses_l2(self):
        """[(variable, l2)] pairs contributing l2 * sum(w^2) to model.losses."""
        return []


class Embedding(Layer):
    """tf.keras.layers.Embedding(input_dim, output_dim): weight [input_dim, output_dim] fp32,
    initializer "uniform" = U(-0.05, 0.05); forward = gather; gradient = IndexedSlices."""

    combiner = None

    def __init__(self, input_dim: int, output_dim: int, name: Optional[str] = None):
        self.input_dim, self.output_dim = int(input_dim), int(output_dim)
        self.name = name or f"embedding_{_layer_counter[0]}"
        g = torch.Generator(device=device())
        g.manual_seed(_next_seed())
        table = torch.empty((self.input_dim, self.output_dim), dtype=torch.float32, device=device())
        table.uniform_(-0.05, 0.05, generator=g)
        self.embeddings = Variable(self.name + "/embeddings", table, "table")

    @property
    def trainable_variables(self):
        return [self.embeddings]

    def get_weights(self):
        return [self.embeddings.numpy()]

    def set_weights(self, weights):
        self.embeddings.assign(weights[0])

    def _feature(self, inputs):
        return (self.embeddings.value, _as_ids(inputs).reshape(-1), None, "sum")

    def _tower_features(self, inputs):
        return [self], [self._feature(inputs)]

    # what the optimizer is told at lookup time / receives as the gradient (overridden by row-sharded tables)
    def _lookup_note(self, feat):
        return (self.embeddings, feat[1], feat[2], feat[3])

    def _make_grad(self, feat, rows) -> IndexedSlices:
        return IndexedSlices(values=feat[1], offsets=feat[2], mode=feat[3], rows=rows)

    def __call__(self, inputs, training: bool = False) -> Tensor:
        return _tower_input([self], [inputs])


class EmbeddingBag(Embedding):
    """Embedding over a ragged id list reduced over the bag axis: reduce_mean / reduce_sum of a
    ragged Embedding output == safe_embedding_lookup_sparse(combiner).  Input: (values, offsets)
    CSR (int64).  Empty bag -> zeros; duplicates inside a bag count each time."""

    def __init__(self, input_dim: int, output_dim: int, combiner: str = "mean", name: Optional[str] = None):
        super().__init__(input_dim, output_dim, name)
        if combiner not in ("sum", "mean"):
            raise ValueError("combiner must be 'sum' or 'mean'")
        self.combiner = combiner

    def _feature(self, inputs):
        values, offsets = inputs
        return (self.embeddings.value, _as_ids(values), _as_ids(offsets), self.combiner)


def _tower_input(layers: Sequence[Embedding], inputs: Sequence) -> Tensor:
    feats = [l._feature(x) for l, x in zip(layers, inputs)]
    f0 = feats[0]
    batch = f0[1].numel() if f0[2] is None else f0[2].numel() - 1
    for f in feats[1:]:
        b = f[1].numel() if f[2] is None else f[2].numel() - 1
        if b != batch:
            raise ValueError(f"features disagree on the batch size ({b} != {batch})")
    dim = layers[0].output_dim
    bf16 = config.precision == "bf16"
    notes = [n for n in (l._lookup_note(f) for l, f in zip(layers, feats)) if n is not None]
    if notes:
        GradientTape.note_sparse_lookup(notes)
    out_f32, out_bf16 = ops.tower_input_fwd(feats, batch, dim, want_f32=not bf16, want_bf16=bf16)
    out = Tensor(f32=out_f32, bf16=out_bf16, grad_formats=("f32",))

    def backward():
        if out.grad is None:
            return
        rows = out.grad["f32"]
        for layer, f in zip(layers, feats):
            if layer.embeddings.grad is not None:
                raise NotImplementedError("an embedding table used twice in one step is not supported")
            layer.embeddings.grad = layer._make_grad(f, rows)

    GradientTape.record(backward)
    return out


class FeatureSum(Layer):
    """Sum of several (pooled) embeddings of one example, fused into a single gather pass:
    out = sum_f pool_f(table_f[ids_f]).  ``features`` maps input-dict keys to Embedding /
    EmbeddingBag layers (cfg3: item id + mean(category) + mean(brand))."""

    def __init__(self, features: Dict[str, Embedding], name: Optional[str] = None):
        if not 1 <= len(features) <= 8:
            raise ValueError("FeatureSum takes 1..8 features")
        dims = {l.output_dim for l in features.values()}
        if len(dims) != 1:
            raise ValueError("all features must share output_dim")
        self.features = dict(features)
        self.name = name or "feature_sum"

    @property
    def trainable_variables(self):
        return [l.embeddings for l in self.features.values()]

    def _tower_features(self, inputs: dict):
        keys = list(self.features)
        layers = [self.features[k] for k in keys]
        return layers, [l._feature(inputs[k]) for l, k in zip(layers, keys)]

    def __call__(self, inputs: dict, training: bool = False) -> Tensor:
        keys = list(self.features)
        return _tower_input([self.features[k] for k in keys], [inputs[k] for k in keys])


class Dense(Layer):
    """tf.keras.layers.Dense(units, activation): kernel [in, out] glorot_uniform, bias zeros,
    optional kernel_regularizer=l2(lam) (lam * sum(w^2) added to the model losses)."""

    def __init__(self, units: int, activation: Optional[str] = None, kernel_regularizer=None,
                 name: Optional[str] = None):
        if activation not in (None, "linear", "relu"):
            raise NotImplementedError(f"activation {activation!r}: only relu / linear are on the hot path")
        self.units = int(units)
        self.relu = activation == "relu"
        self.l2 = float(getattr(kernel_regularizer, "l2", kernel_regularizer) or 0.0)
        self.name = name or f"dense_{_layer_counter[0]}"
        self._seed = _next_seed()
        self.kernel: Optional[Variable] = None
        self.bias: Optional[Variable] = None

    def build(self, in_dim: int) -> None:
        rng = np.random.Generator(np.random.PCG64(self._seed))
        lim = math.sqrt(6.0 / (in_dim + self.units))
        k = rng.uniform(-lim, lim, size=(in_dim, self.units)).astype(np.float32)
        dev = device()
        self.kernel = Variable(self.name + "/kernel", torch.from_numpy(k).to(dev), "kernel", l2=self.l2)
        self.bias = Variable(self.name + "/bias", torch.zeros(self.units, dtype=torch.float32, device=dev), "bias")
        if config.precision == "bf16":
            self.kernel.want_shadows = True
            self.kernel.refresh_shadows()

    @property
    def trainable_variables(self):
        return [] if self.kernel is None else [self.kernel, self.bias]

    @property
    def losses_l2(self):
        return [(self.kernel, self.l2)] if (self.kernel is not None and self.l2) else []

    def get_weights(self):
        return [self.kernel.numpy(), self.bias.numpy()]

    def set_weights(self, weights):
        if self.kernel is None:
            self.build(np.asarray(weights[0]).shape[0])
        self.kernel.assign(weights[0])
        self.bias.assign(weights[1])

    def __call__(self, x: Tensor, training: bool = False) -> Tensor:
        in_dim = x.shape[1]
        if self.kernel is None:
            self.build(in_dim)
        prec = config.precision
        if prec == "fp32":
            y, _ = ops.dense_fwd("fp32", x.f32, self.kernel.value, self.bias.value, self.relu)
            out = Tensor(f32=y, grad_formats=("f32",))
        else:
            if self.kernel.shadow is None:
                self.kernel.want_shadows = True
                self.kernel.refresh_shadows()
            y, _ = ops.dense_fwd("bf16", x.bf16, self.kernel.shadow, self.bias.value, self.relu)
            out = Tensor(bf16=y, grad_formats=("bf16",))
        out.relu_output = self.relu

        def backward():
            if out.grad is None:
                return
            g = out.grad
            want = x.grad_formats
            need_dx = x.producer_needs_grad
            if prec == "fp32":
                dx, _, dk, P, db = ops.dense_bwd("fp32", g["f32"], x.f32, self.kernel.value,
                                                 relu_mask_x=x.relu_output, want_dx=need_dx)
                if need_dx:
                    x.grad = dict(f32=dx)
            else:
                dy_f32 = g.get("f32")
                dx, dx_f32, dk, P, db = ops.dense_bwd(
                    "bf16", g["bf16"], x.bf16, self.kernel.shadow, relu_mask_x=x.relu_output,
                    want_dx=need_dx and "bf16" in want,
                    want_dx_f32=need_dx and ("f32" in want or "parts" in want),
                    want_dbias=dy_f32 is None)
                if need_dx:
                    x.grad = dict(f32=dx_f32, bf16=dx)
                if dy_f32 is not None:      # bias gradient from the un-rounded upstream gradient
                    db = ops.colsum_f32(dy_f32)
            self.kernel.grad = DenseGrad(dk, P)
            db = db.reshape(-1, 1, self.units)
            self.bias.grad = DenseGrad(db, db.shape[0])

        GradientTape.record(backward)
        return out


class PendingTowers:
    """Tower outputs that have been requested but not computed yet.  The first access to any of
    them runs ONE fused launch (ops.tower_mlp2_fwd: gather + pool + Dense(relu) + Dense for every
    queued tower, query and candidate side together) and records ONE backward node that will run
    ops.tower_mlp2_bwd for all of them."""

    def __init__(self):
        self.items = []

    def add(self, seq, emb_layers, feats, batch, out: Tensor, x_input: Optional[Tensor] = None):
        """x_input: the tower input as a Tensor (bf16) produced elsewhere -- a row-sharded lookup that has
        already exchanged its rows -- instead of features to gather; its gradient is handed back as f32."""
        self.items.append((seq, emb_layers, feats, batch, out, x_input))

    def flush(self):
        items, self.items = self.items, []
        if not items:
            return
        lookups = [l._lookup_note(f) for _, emb_layers, feats, _, _, _ in items for l, f in zip(emb_layers, feats)]
        lookups = [n for n in lookups if n is not None]       # None: announced later (peer-memory exchange)
        # ID-only local towers: the tower kernel's control warp dedups the ids itself (no launch, no stream join)
        def id_only_local(emb_layers, feats, x_input):
            return (x_input is None and len(feats) == 1 and len(emb_layers) == 1 and feats[0][2] is None
                    and not isinstance(feats[0][0], ops.PeerTable) and emb_layers[0]._lookup_note(feats[0]) is not None)
        fused_prepare = Sequential.fuse_prepare and bool(lookups) and all(id_only_local(e, f, xi) for _, e, f, _, _, xi in items)
        prepare_ws = None
        if lookups:
            if fused_prepare:
                prepare_ws = GradientTape.note_sparse_lookup(lookups, fused=True)
            else:
                GradientTape.note_sparse_lookup(lookups)
        specs = []
        for k, (seq, emb_layers, feats, batch, out, x_input) in enumerate(items):
            d1, d2 = seq.layers[1], seq.layers[2]
            spec = dict(features=feats, batch=batch, w1=d1.kernel.shadow, b1=d1.bias.value,
                        w2=d2.kernel.shadow, b2=d2.bias.value)
            if x_input is not None:
                spec["x_input"] = x_input.bf16
            elif Sequential.pregather_bags and (len(feats) > 1 or any(f[2] is not None for f in feats)) \
                    and not any(isinstance(f[0], ops.PeerTable) for f in feats):
                # Several rows per example (bags, several features): the stand-alone gather kernel -- a warp per output
                # row, every SM full of them -- pools the tower input at 3 TB/s, where the 128 CTAs of the tower kernel
                # keep 8 gathering warps each (cfg3 item tower, 7 rows per example: 21.6 us + a TMA-fed tower kernel
                # against 86 us for the in-kernel gather).  One row per example stays in the tower kernel.
                _, xb = ops.tower_input_fwd(feats, batch, d1.kernel.shape[0], want_f32=False, want_bf16=True)
                spec["features"] = []
                spec["x_input"] = xb
            if prepare_ws is not None:
                spec["prepare_ws"] = prepare_ws[k]
            specs.append(spec)
        results = ops.tower_mlp2_fwd(specs)
        for (seq, emb_layers, feats, batch, out, x_input), spec, (x, h, y) in zip(items, specs, results):
            out._bf16 = y
            out._pending = None
            spec["x"], spec["h"] = x, h

        def backward():
            live = [(it, sp) for it, sp in zip(items, specs) if it[4].grad is not None]
            for _, sp in live:
                sp.pop("x_input", None)
            if not live:
                return
            towers = []
            for (seq, emb_layers, feats, batch, out, x_input), sp in live:
                g = out.grad
                parts = g["parts"] if g.get("parts") is not None else g["f32"].reshape(1, *g["f32"].shape)
                towers.append(dict(sp, dy_parts=parts.contiguous(), dy_splits=parts.shape[0]))
            outs = ops.tower_mlp2_bwd(towers)
            for ((seq, emb_layers, feats, batch, out, x_input), sp), o in zip(live, outs):
                P = o["P"]
                if x_input is not None:
                    x_input.grad = dict(f32=o["dx"])
                for layer, f in zip(emb_layers, feats):
                    if layer.embeddings.grad is not None:
                        raise NotImplementedError("an embedding table used twice in one step is not supported")
                    layer.embeddings.grad = layer._make_grad(f, o["dx"])
                d1, d2 = seq.layers[1], seq.layers[2]
                d1.kernel.grad = DenseGrad(o["dw1"], P)
                d1.bias.grad = DenseGrad(o["db1"].reshape(P, 1, -1), P)
                d2.kernel.grad = DenseGrad(o["dw2"], P)
                d2.bias.grad = DenseGrad(o["db2"].reshape(P, 1, -1), P)

        GradientTape.record(backward)


_pending_towers = PendingTowers()


class Sequential(Layer):
    """tf.keras.Sequential over the layers above.  In bf16 precision a [Embedding | FeatureSum,
    Dense(relu), Dense] tower of a supported shape (ops.tower_mlp2_supported) runs as the fused
    tower kernels; anything else runs layer by layer."""

    fuse = True
    # id dedup by a dedicated warp of the tower forward (ID-only towers) instead of tt_optimizer_prepare_sparse on a side
    # stream.  OFF by default: under the memory load of the gather a dependent round trip costs 1.3-2 us, the insert
    # needs >= 3 of them per id (id -> key line -> CAS), and the tower kernel waits for its slowest chain: measured at
    # cfg2 the forward grows by 5.6 us while the removed stream join saves 2.9 us.
    fuse_prepare = os.environ.get("TT_FUSED_PREPARE", "0") == "1"
    # towers with bags / several features: pool the tower input with the stand-alone gather kernel ahead of the fused MLP launch
    pregather_bags = os.environ.get("TT_PREGATHER", "1") == "1"

    def __init__(self, layers: Sequence[Layer] = (), name: Optional[str] = None):
        self.layers = list(layers)
        self.name = name or "sequential"

    def add(self, layer: Layer) -> None:
        self.layers.append(layer)

    @property
    def trainable_variables(self):
        return [v for l in self.layers for v in l.trainable_variables]

    @property
    def losses_l2(self):
        return [p for l in self.layers for p in l.losses_l2]

    def _fusable(self) -> bool:
        if not (self.fuse and config.precision == "bf16" and len(self.layers) == 3):
            return False
        e, d1, d2 = self.layers
        sharded = getattr(e, "provides_tower_input", False)      # parallel.ShardedEmbedding
        gathers = isinstance(e, (Embedding, FeatureSum)) or hasattr(e, "_tower_features")   # incl. PeerShardedEmbedding
        if not (gathers or sharded) or not isinstance(d1, Dense) or not isinstance(d2, Dense):
            return False
        if not d1.relu or d2.relu:
            return False
        d_in = next(iter(e.features.values())).output_dim if isinstance(e, FeatureSum) else e.output_dim
        return ops.tower_mlp2_supported(d_in, d1.units, d2.units)

    def __call__(self, inputs, training: bool = False) -> Tensor:
        if self._fusable() and getattr(self.layers[0], "provides_tower_input", False):
            e, d1, d2 = self.layers
            xt = e(inputs, training=training)                     # exchanges the rows; records its own backward
            batch = xt.shape[0]
            if batch > 0:
                for layer, width in ((d1, e.output_dim), (d2, d1.units)):
                    if layer.kernel is None:
                        layer.build(width)
                    if layer.kernel.shadow is None:
                        layer.kernel.want_shadows = True
                        layer.kernel.refresh_shadows()
                out = Tensor(grad_formats=("parts",), pending=_pending_towers, shape=(batch, d2.units))
                _pending_towers.add(self, [], [], batch, out, x_input=xt)
                return out
            x = xt
            for l in self.layers[1:]:
                x = l(x, training=training)
            return x
        if self._fusable():
            e, d1, d2 = self.layers
            emb_layers, feats = e._tower_features(inputs)
            batch = feats[0][1].numel() if feats[0][2] is None else feats[0][2].numel() - 1
            for f in feats[1:]:
                b = f[1].numel() if f[2] is None else f[2].numel() - 1
                if b != batch:
                    raise ValueError(f"features disagree on the batch size ({b} != {batch})")
            if batch > 0:
                d_in = emb_layers[0].output_dim
                for layer, width in ((d1, d_in), (d2, d1.units)):
                    if layer.kernel is None:
                        layer.build(width)
                    if layer.kernel.shadow is None:
                        layer.kernel.want_shadows = True
                        layer.kernel.refresh_shadows()
                out = Tensor(grad_formats=("parts",), pending=_pending_towers, shape=(batch, d2.units))
                _pending_towers.add(self, emb_layers, feats, batch, out)
                return out
        x = inputs
        for l in self.layers:
            x = l(x, training=training)
        return x


from . import factorized_top_k  # noqa: E402,F401
