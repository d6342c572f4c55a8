"""tfrs.models.Model: user overrides compute_loss; train_step runs the tape,
adds the regularization losses, takes gradients and applies them."""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import numpy as np
import torch

from . import ops
from .core import GradientTape, Scalar, Variable
from .layers import Layer


class Model:
    def __init__(self, name: Optional[str] = None):
        self.name = name or type(self).__name__
        self.optimizer = None

    # ---- Keras-ish plumbing ----------------------------------------------------------
    def compile(self, optimizer=None, **_ignored) -> None:
        self.optimizer = optimizer

    def _sublayers(self) -> List[Layer]:
        seen, out = set(), []

        def visit(obj):
            if id(obj) in seen:
                return
            seen.add(id(obj))
            if isinstance(obj, Layer):
                out.append(obj)
                return
            if isinstance(obj, (list, tuple)):
                for o in obj:
                    visit(o)
            elif isinstance(obj, dict):
                for o in obj.values():
                    visit(o)

        for v in self.__dict__.values():
            visit(v)
        return out

    @property
    def trainable_variables(self) -> List[Variable]:
        seen, out = set(), []
        for l in self._sublayers():
            for v in l.trainable_variables:
                if id(v) not in seen:
                    seen.add(id(v))
                    out.append(v)
        return out

    @property
    def metrics(self):
        ms = []
        for v in self.__dict__.values():
            fm = getattr(v, "factorized_metrics", None)
            if fm is not None:
                ms.append(fm)
        return ms

    def _regularization_loss(self) -> Optional[torch.Tensor]:
        pairs = [p for l in self._sublayers() for p in l.losses_l2]
        if not pairs:
            return None
        out = torch.empty(1, dtype=torch.float32, device=pairs[0][0].value.device)
        for i, (var, lam) in enumerate(pairs):
            ops.sum_squares(var.value, lam, out, accumulate=i > 0)
        return out

    # ---- the TFRS contract -----------------------------------------------------------
    def compute_loss(self, inputs, training: bool = False) -> Scalar:
        raise NotImplementedError("Implementers must implement the `compute_loss` method.")

    def train_step(self, inputs) -> Dict[str, object]:
        if self.optimizer is None:
            raise RuntimeError("call model.compile(optimizer=...) before train_step")
        self.optimizer.begin_step()
        with GradientTape() as tape:
            tape.on_sparse_lookup = self.optimizer.prepare_sparse
            loss = self.compute_loss(inputs, training=True)
            self.optimizer.join_prepare()
            reg = self._regularization_loss()
            variables = self.trainable_variables
            grads = tape.gradient(loss, variables)
        self.optimizer.apply_gradients(zip(grads, variables))
        return self._report(loss, reg)

    def test_step(self, inputs) -> Dict[str, object]:
        loss = self.compute_loss(inputs, training=False)
        return self._report(loss, self._regularization_loss())

    def _report(self, loss: Scalar, reg: Optional[torch.Tensor]) -> Dict[str, object]:
        out = {}
        for m in self.metrics:
            out.update(m.result())
        zero = torch.zeros_like(loss.value) if reg is None else reg
        out["loss"] = loss.value
        out["regularization_loss"] = zero
        out["total_loss"] = loss.value if reg is None else loss.value + reg
        return out

    # ---- raw-tensor checkpoint: tables, Dense kernels / biases, optimizer slots -----------------
    def save_weights(self, path) -> None:
        """One .npz: every trainable variable (fp32 master copy) in `trainable_variables` order, its optimizer slots
        (Adagrad accumulator / Adam m, v) and the optimizer's iteration count.  Layers must be built."""
        arrays, names = {}, []
        for i, v in enumerate(self.trainable_variables):
            names.append(v.name)
            arrays[f"v{i}"] = v.value.detach().cpu().numpy()
            for k, s in v.slots.items():
                if isinstance(s, torch.Tensor) and not k.startswith("_"):
                    arrays[f"s{i}.{k}"] = s.detach().cpu().numpy()
        arrays["names"] = np.array(names)
        arrays["iterations"] = np.array([self.optimizer.iterations if self.optimizer is not None else 0], dtype=np.int64)
        with open(path, "wb") as f:
            np.savez(f, **arrays)

    def load_weights(self, path) -> None:
        """Inverse of save_weights onto a BUILT model of the same architecture (shapes are checked)."""
        z = np.load(path, allow_pickle=False)
        variables = self.trainable_variables
        n = len(z["names"])
        if n != len(variables):
            raise ValueError(f"checkpoint holds {n} variables, the model has {len(variables)}")
        for i, v in enumerate(variables):
            a = z[f"v{i}"]
            if tuple(a.shape) != v.shape:
                raise ValueError(f"variable {i} ({v.name}): checkpoint shape {tuple(a.shape)} != {v.shape}")
            v.assign(a)                       # in place: value and bf16 shadow keep their addresses (captured graphs)
            prefix = f"s{i}."
            for key in z.files:
                if key.startswith(prefix):
                    v.assign_slot(key[len(prefix):], z[key])
        if self.optimizer is not None:
            self.optimizer.set_iterations(int(z["iterations"][0]))

    def fit(self, dataset: Iterable, epochs: int = 1, verbose: int = 0):
        history = {"loss": [], "total_loss": []}
        for _ in range(epochs):
            last = None
            for batch in dataset:
                last = self.train_step(batch)
            if last is not None:
                history["loss"].append(float(last["loss"].item()))
                history["total_loss"].append(float(last["total_loss"].item()))
        return history

    def evaluate(self, dataset: Iterable, return_dict: bool = True):
        for m in self.metrics:
            m.reset_states()
        last = None
        for batch in dataset:
            last = self.test_step(batch)
        if last is None:
            return {}
        return {k: (float(v.item()) if isinstance(v, torch.Tensor) else v) for k, v in last.items()}

    # ---- CUDA-graph replay of the whole step (launch-bound at cfg2) ------------------
    def make_graphed_train_step(self, example_inputs: Dict[str, torch.Tensor], warmup: int = 3,
                                steps_per_execution: int = 1):
        """Capture train_step into a CUDA graph with static input buffers.  Returns
        step(inputs) -> dict of device scalars (valid until the next replay).
        steps_per_execution = S > 1 (the Keras `Model.compile(steps_per_execution=S)` idea): one graph holds S
        consecutive train steps; step(batches) takes a sequence of S batches (one H2D copy for all of them) and
        returns the S result dicts.  Inside the graph step i + 1 follows step i as a programmatic dependent launch
        (~1.5 us) instead of across a graph-launch boundary (~10 us at cfg2)."""
        return GraphedStep(self, example_inputs, warmup, steps_per_execution=steps_per_execution)


class GraphedStep:
    """CUDA-graph replay of `model.train_step` with DOUBLE-BUFFERED static inputs: the step is captured `buffers` times,
    each capture reading its own packed input buffer, and successive calls alternate between them.  The copy of step
    i + 1's inputs (one H2D copy from a pinned staging ring for host batches, one D2D copy for `pack`ed device batches)
    is issued on a copy stream and overlaps the replay of step i; the replay only waits for its own inputs."""

    def __init__(self, model: Model, example_inputs, warmup: int, buffers: int = 2, steps_per_execution: int = 1):
        self.model = model
        self.steps_per_execution = S = max(1, int(steps_per_execution))
        side = torch.cuda.Stream()
        self._slots = []
        # S > 1: the static inputs are a tuple of S batches in ONE flat buffer (one copy per execution)
        example = example_inputs if S == 1 else tuple(example_inputs for _ in range(S))
        first = _PackedInputs(example)
        views = lambda packed: (packed.device_views,) if S == 1 else packed.device_views
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                model.train_step(views(first)[0])
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        iterations = model.optimizer.iterations
        for b in range(max(1, buffers)):
            packed = first if b == 0 else _PackedInputs(example)
            graph = torch.cuda.CUDAGraph()
            before = ops.LAUNCHES
            # the S losses of an execution land in ONE tensor (the first retrieval loss of each step): `losses`
            losses = torch.zeros(S, dtype=torch.float32, device=packed.flat_dev.device)
            free = []
            ops.loss_allocator = lambda: free.pop() if free else None
            try:
                with torch.cuda.graph(graph):
                    outs = []
                    for k, v in enumerate(views(packed)):
                        free[:] = [losses[k:k + 1]]
                        outs.append(model.train_step(v))
            finally:
                ops.loss_allocator = None
            self.launches_per_replay = ops.LAUNCHES - before
            slot = _Slot(packed, graph, outs[0] if S == 1 else outs)
            slot.losses = losses
            self._slots.append(slot)
        model.optimizer.iterations = iterations       # the captures executed nothing: no step was taken
        self._copy_stream = torch.cuda.Stream()
        self._next = 0
        # single-buffer compatibility names (slot 0)
        self.static_in, self.static_out, self.graph = first.device_views, self._slots[0].out, self._slots[0].graph

    def pack(self, inputs) -> "PackedBatch":
        """A device-resident batch laid out like the static inputs: `step(packed)` then costs one D2D copy."""
        return self._slots[0].packed.pack(inputs)

    def __call__(self, inputs=None):
        """inputs: one batch (steps_per_execution == 1) or a sequence of steps_per_execution batches -- host tensors,
        device tensors, or a PackedBatch from pack(); None replays on the inputs already in place."""
        if self.steps_per_execution > 1 and inputs is not None and not isinstance(inputs, PackedBatch):
            inputs = tuple(inputs)
            if len(inputs) != self.steps_per_execution:
                raise ValueError(f"expected {self.steps_per_execution} batches per execution, got {len(inputs)}")
        slot = self._slots[self._next]
        self._next = (self._next + 1) % len(self._slots)
        main = torch.cuda.current_stream()
        if inputs is not None:
            cs = self._copy_stream
            cs.wait_event(slot.consumed)              # the replay that last read this slot's inputs has finished
            with torch.cuda.stream(cs):
                if isinstance(inputs, PackedBatch):
                    slot.packed.flat_dev.copy_(inputs.flat, non_blocking=True)
                elif not slot.packed.load_from_host(inputs):
                    # device tensors of unknown provenance: whatever produced them was enqueued on the caller's stream,
                    # so this copy has to run behind it (no overlap with the previous replay; host batches and
                    # pack()ed batches do overlap)
                    cs.wait_stream(main)
                    _copy_inputs(slot.packed.device_views, inputs)
                slot.ready.record(cs)
            main.wait_event(slot.ready)
        slot.graph.replay()
        slot.consumed.record(main)
        ops._count(self.launches_per_replay)
        self.model.optimizer.iterations += self.steps_per_execution   # host mirror of the step count (Adam's device counter advanced in-graph)
        self.losses = slot.losses                     # device [S]: the retrieval losses of this execution's steps, contiguous
        return slot.out


class _Slot:
    def __init__(self, packed, graph, out):
        self.packed, self.graph, self.out = packed, graph, out
        self.ready, self.consumed = torch.cuda.Event(), torch.cuda.Event()
        self.consumed.record()


class PackedBatch:
    """Device-resident batch in the flat layout of a GraphedStep's static inputs (GraphedStep.pack)."""

    def __init__(self, flat):
        self.flat = flat


class _PackedInputs:
    """All tensors of an input structure as views of one flat device buffer + one pinned host staging buffer."""

    def __init__(self, example):
        leaves = []
        _leaves(example, leaves)
        dev = next((t.device for t in leaves if t.is_cuda), None)
        if dev is None:                               # host example (e.g. data.InteractionBatches.example()): static inputs live on the GPU
            dev = torch.device("cuda", torch.cuda.current_device())
        offs, off = [], 0
        for t in leaves:
            offs.append(off)
            off += (t.numel() * t.element_size() + 255) // 256 * 256
        self.flat_dev = torch.zeros(max(off, 256), dtype=torch.uint8, device=dev)
        mk = lambda flat, t, o: flat[o:o + t.numel() * t.element_size()].view(t.dtype).view(t.shape)
        dviews = [mk(self.flat_dev, t, o) for t, o in zip(leaves, offs)]
        # a small ring of pinned staging buffers: slot i is rewritten only after the H2D copy that read it has finished
        self._ring = []
        for _ in range(4):
            fh = torch.zeros(max(off, 256), dtype=torch.uint8).pin_memory()
            self._ring.append((fh, [mk(fh, t, o) for t, o in zip(leaves, offs)], torch.cuda.Event()))
        self._next = 0
        self._n_leaves = len(leaves)
        self._offs, self._nbytes = offs, [t.numel() * t.element_size() for t in leaves]
        for v, t in zip(dviews, leaves):
            v.copy_(t.to(dev))
        it = iter(dviews)
        self.device_views = _rebuild(example, it)

    def pack(self, inputs) -> "PackedBatch":
        leaves = []
        _leaves(inputs, leaves)
        if len(leaves) != self._n_leaves:
            raise ValueError("pack: the batch does not have the structure of the example inputs")
        flat = torch.zeros_like(self.flat_dev)
        for t, o, n in zip(leaves, self._offs, self._nbytes):
            if t.numel() * t.element_size() != n:
                raise ValueError("pack: tensor sizes differ from the example inputs (static shapes)")
            flat[o:o + n].copy_(t.to(flat.device).contiguous().view(-1).view(torch.uint8))
        torch.cuda.current_stream().synchronize()     # complete before any copy stream may read it
        return PackedBatch(flat)

    def load_from_host(self, inputs) -> bool:
        """If every tensor of `inputs` lives in host memory: stage them and issue ONE H2D copy.  Else False."""
        leaves = []
        _leaves(inputs, leaves)
        if len(leaves) != self._n_leaves or any(t.device.type != "cpu" for t in leaves):
            return False
        flat_host, hviews, ev = self._ring[self._next]
        self._next = (self._next + 1) % len(self._ring)
        ev.synchronize()
        for h, t in zip(hviews, leaves):
            h.copy_(t)
        self.flat_dev.copy_(flat_host, non_blocking=True)
        ev.record()
        return True


def _leaves(x, out):
    if isinstance(x, torch.Tensor):
        out.append(x)
    elif isinstance(x, dict):
        for k in x:
            _leaves(x[k], out)
    elif isinstance(x, (tuple, list)):
        for v in x:
            _leaves(v, out)


def _rebuild(x, it):
    if isinstance(x, torch.Tensor):
        return next(it)
    if isinstance(x, dict):
        return {k: _rebuild(v, it) for k, v in x.items()}
    if isinstance(x, (tuple, list)):
        return type(x)(_rebuild(v, it) for v in x)
    return x


def _clone_inputs(x):
    if isinstance(x, torch.Tensor):
        return x.clone()
    if isinstance(x, dict):
        return {k: _clone_inputs(v) for k, v in x.items()}
    if isinstance(x, (tuple, list)):
        return type(x)(_clone_inputs(v) for v in x)
    return x


def _copy_inputs(dst, src):
    if isinstance(dst, torch.Tensor):
        dst.copy_(src, non_blocking=True)
    elif isinstance(dst, dict):
        for k in dst:
            _copy_inputs(dst[k], src[k])
    elif isinstance(dst, (tuple, list)):
        for d, s in zip(dst, src):
            _copy_inputs(d, s)
