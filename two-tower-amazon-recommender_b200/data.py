"""Input-format bridge: the interaction frames the reference's data layer writes ->
int64 id batches for `Model.train_step` / `GraphedStep`.

Two on-disk schemas exist in the reference and both are accepted:

* `scripts/data_processing/prepare_training_data.py:204-234` writes `combined_interactions.parquet` with the dense
  ids `user_idx` / `item_idx` (next to `user_id`, `parent_asin`, `category`, `rating`, `timestamp`) and a
  `mappings.pkl` holding `user_to_idx`, `item_to_idx` and their inverses;
* `src/data/preprocessor.py:478-491` (`_encode_categories`) adds `user_id_encoded`, `item_id_encoded` and, when the
  frame has a `main_category` column, `category_encoded`.

Every batch is a dict of int64 tensors under the file's own column names AND under the canonical keys the model code
of INTEGRATION.md uses (`user_id_encoded`, `item_id_encoded`, `category_encoded`).  Host batches come out of a small
ring of (optionally pinned) staging tensors, which is what `GraphedStep.__call__` turns into one H2D copy per step;
`to_device()` keeps the id columns resident on the GPU and draws every batch there (a device permutation per epoch +
`index_select`: no per-step host traffic at all).  No feature engineering, text or image handling lives here: that is
the reference's ETL and out of scope.
"""
from __future__ import annotations

import pickle
from pathlib import Path
from typing import Dict, Iterator, Optional, Sequence, Union

import numpy as np
import torch

# (column in the file, canonical key) per schema, in detection order
_SCHEMAS = (
    (("user_id_encoded", "user_id_encoded"), ("item_id_encoded", "item_id_encoded")),      # preprocessor.py:481-482
    (("user_idx", "user_id_encoded"), ("item_idx", "item_id_encoded")),                    # prepare_training_data.py:209-210
)
_OPTIONAL = (("category_encoded", "category_encoded"),)                                    # preprocessor.py:487


def load_mappings(path: Union[str, Path]) -> Dict[str, dict]:
    """`mappings.pkl` of prepare_training_data.py:222-232: user_to_idx / item_to_idx / idx_to_user / idx_to_item."""
    with open(path, "rb") as f:
        m = pickle.load(f)  # nosec B301 - the file is produced by the reference's own pipeline
    for k in ("user_to_idx", "item_to_idx"):
        if k not in m:
            raise ValueError(f"{path}: not a mappings.pkl of the reference (missing {k!r})")
    return m


def _read_columns(source, wanted: Sequence[str]) -> Dict[str, np.ndarray]:
    """Columns of a parquet file / pandas frame / dict of arrays as contiguous int64 arrays (missing ones skipped)."""
    out = {}
    if isinstance(source, (str, Path)):
        import pyarrow.parquet as pq
        names = set(pq.ParquetFile(str(source)).schema_arrow.names)
        cols = [c for c in wanted if c in names]
        table = pq.read_table(str(source), columns=cols)
        for c in cols:
            out[c] = table.column(c).to_numpy(zero_copy_only=False)
    else:
        for c in wanted:
            if c in source:
                out[c] = np.asarray(source[c])
    for c, v in out.items():
        if v.dtype.kind not in "iu":
            raise TypeError(f"column {c!r} must hold integer ids (got {v.dtype}); run the reference's encoding step first")
        out[c] = np.ascontiguousarray(v, dtype=np.int64)
    return out


class InteractionBatches:
    """Epoch iterator over the interactions of one frame.  A yielded batch stays valid until `ring` further batches
    have been drawn AND the device work the consumer enqueued for it (on the current stream) has finished.

        ds = InteractionBatches("data/processed/combined_interactions.parquet", batch_size=8192, seed=0)
        model = TwoTower(ds.num_users, ds.num_items); step = model.make_graphed_train_step(ds.example())
        for epoch in range(E):
            for batch in ds:            # dict of pinned int64 tensors, valid until `ring` batches later
                step(batch)

    shuffle: a fresh permutation per epoch from numpy's PCG64(seed + epoch) (the generator the synthetic configs use);
    drop_remainder: keep every batch the same size (what a captured CUDA graph needs)."""

    def __init__(self, source, batch_size: int, shuffle: bool = True, seed: int = 0, drop_remainder: bool = True,
                 pin: Optional[bool] = None, ring: int = 4):
        wanted = [c for schema in _SCHEMAS for c, _ in schema] + [c for c, _ in _OPTIONAL]
        cols = _read_columns(source, wanted)
        schema = next((s for s in _SCHEMAS if all(c in cols for c, _ in s)), None)
        if schema is None:
            raise ValueError("no id columns found: expected user_id_encoded/item_id_encoded (preprocessor.py:481-482) or "
                             "user_idx/item_idx (prepare_training_data.py:209-210)")
        self._columns = {}                     # output key -> int64 array [n]
        for c, canon in tuple(schema) + tuple(p for p in _OPTIONAL if p[0] in cols):
            self._columns[c] = cols[c]
            self._columns[canon] = cols[c]
        self.user_key, self.item_key = schema[0][0], schema[1][0]
        n = len(cols[self.user_key])
        if any(len(v) != n for v in self._columns.values()):
            raise ValueError("id columns disagree on the number of rows")
        for c, v in self._columns.items():
            if n and v.min() < 0:
                raise ValueError(f"column {c!r} holds negative ids")
        self.num_rows = n
        self.num_users = int(cols[self.user_key].max()) + 1 if n else 0
        self.num_items = int(cols[self.item_key].max()) + 1 if n else 0
        self.num_categories = int(cols["category_encoded"].max()) + 1 if "category_encoded" in cols and n else 0
        self.batch_size = int(batch_size)
        if self.batch_size <= 0:
            raise ValueError("batch_size must be positive")
        self.shuffle, self.seed, self.drop_remainder = bool(shuffle), int(seed), bool(drop_remainder)
        self.epoch = 0
        self._pin = torch.cuda.is_available() if pin is None else bool(pin)
        self._ring = [self._alloc() for _ in range(max(2, int(ring)))]
        # one CUDA event per staging slot: recorded after the consumer's turn, waited for before the slot is rewritten
        # (the consumer may copy a pinned batch to the device asynchronously and run many steps ahead of the GPU)
        self._events = [torch.cuda.Event() for _ in self._ring] if self._pin and torch.cuda.is_available() else None
        self._slot = 0

    def _alloc(self):
        out = {}
        done = {}
        for k, v in self._columns.items():
            if id(v) not in done:              # canonical aliases share the staging tensor of their column
                t = torch.empty(self.batch_size, dtype=torch.int64)
                done[id(v)] = t.pin_memory() if self._pin else t
            out[k] = done[id(v)]
        return out

    def __len__(self) -> int:
        return self.num_rows // self.batch_size if self.drop_remainder else -(-self.num_rows // self.batch_size)

    def example(self) -> Dict[str, torch.Tensor]:
        """A batch-shaped dict (first rows, unshuffled) for building layers / capturing a graph."""
        return {k: torch.from_numpy(v[:self.batch_size].copy()) for k, v in self._columns.items()}

    def order(self, epoch: int) -> np.ndarray:
        if not self.shuffle:
            return np.arange(self.num_rows, dtype=np.int64)
        return np.random.Generator(np.random.PCG64(self.seed + epoch)).permutation(self.num_rows).astype(np.int64)

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        perm = self.order(self.epoch)
        self.epoch += 1
        for b in range(len(self)):
            rows = perm[b * self.batch_size:(b + 1) * self.batch_size]
            slot = self._slot
            stage = self._ring[slot]
            self._slot = (self._slot + 1) % len(self._ring)
            if self._events is not None:
                self._events[slot].synchronize()
            seen = set()
            for k, v in self._columns.items():
                t = stage[k]
                if id(t) in seen:
                    continue
                seen.add(id(t))
                np.take(v, rows, out=t.numpy()[:len(rows)])
            yield stage if len(rows) == self.batch_size else {k: t[:len(rows)] for k, t in stage.items()}
            if self._events is not None:
                self._events[slot].record()          # everything the consumer enqueued for this batch precedes it

    def to_device(self, device=None) -> "DeviceInteractionBatches":
        return DeviceInteractionBatches(self, device)


class DeviceInteractionBatches:
    """The same epochs with the id columns resident in HBM (16 bytes per interaction for the two id columns): every
    batch is drawn on the device by a per-epoch device permutation + index_select, so a step reads nothing from the
    host.  The permutation comes from torch's device generator seeded with seed + epoch (not the host PCG64 order)."""

    def __init__(self, host: InteractionBatches, device=None):
        self.host = host
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        done = {}
        self._columns = {}
        for k, v in host._columns.items():
            if id(v) not in done:
                done[id(v)] = torch.tensor(v, dtype=torch.int64, device=self.device)
            self._columns[k] = done[id(v)]
        self.epoch = 0

    def __len__(self) -> int:
        return len(self.host)

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        h = self.host
        if h.shuffle:
            g = torch.Generator(device=self.device)
            g.manual_seed(h.seed + self.epoch)
            perm = torch.randperm(h.num_rows, device=self.device, generator=g)
        else:
            perm = torch.arange(h.num_rows, device=self.device)
        self.epoch += 1
        for b in range(len(self)):
            rows = perm[b * h.batch_size:(b + 1) * h.batch_size]
            out, seen = {}, {}
            for k, col in self._columns.items():
                if id(col) not in seen:
                    seen[id(col)] = col.index_select(0, rows)
                out[k] = seen[id(col)]
            yield out
