"""tfrs.layers.factorized_top_k: BruteForce / Streaming / TopK.  Exact
retrieval: scores = q @ cand^T and top-k fused in one libtwotower kernel, never materialising
the score matrix.  Result order is tf.math.top_k's: score descending, ties -> lower index."""
from __future__ import annotations

from typing import Iterable, Optional

import numpy as np
import torch

from .. import ops
from ..core import Tensor, config, device


def _to_device_matrix(x, precision: str) -> torch.Tensor:
    if isinstance(x, Tensor):
        if precision == "bf16":
            return x.bf16 if x.bf16 is not None else ops.cast_f32_to_bf16(x.f32)
        return x.f32 if x.f32 is not None else x.bf16.float()
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
    if not t.is_cuda:
        t = t.to(device())
    if precision == "bf16":
        if t.dtype == torch.bfloat16:
            return t.contiguous()
        return ops.cast_f32_to_bf16(t.float().contiguous())
    return t.float().contiguous()


class TopK:
    """Interface of tfrs.layers.factorized_top_k.TopK."""

    def __init__(self, k: int = 10, name: Optional[str] = None):
        self._k = int(k)
        self.name = name or type(self).__name__

    def is_exact(self) -> bool:
        raise NotImplementedError

    def index_from_dataset(self, candidates: Iterable):
        """candidates yields embeddings [n,d] or (identifiers [n], embeddings [n,d]) batches."""
        embs, ids = [], []
        for batch in candidates:
            if isinstance(batch, (tuple, list)) and len(batch) == 2:
                ids.append(torch.as_tensor(np.asarray(batch[0])) if not isinstance(batch[0], torch.Tensor) else batch[0])
                e = batch[1]
            else:
                e = batch
            if isinstance(e, Tensor):
                e = e.torch()
            embs.append(e if isinstance(e, torch.Tensor) else torch.as_tensor(np.asarray(e)))
        if not embs:
            raise ValueError("index_from_dataset: empty candidate dataset")
        dev = device()
        cands = torch.cat([e.to(dev) for e in embs], dim=0)
        idents = torch.cat([i.to(dev) for i in ids], dim=0) if ids else None
        return self.index(cands, idents)


class BruteForce(TopK):
    """tfrs.layers.factorized_top_k.BruteForce(query_model=None, k=10)."""

    def __init__(self, query_model=None, k: int = 10, name: Optional[str] = None, precision: Optional[str] = None):
        super().__init__(k, name)
        self.query_model = query_model
        self.precision = precision
        self._candidates = None
        self._identifiers = None

    def is_exact(self) -> bool:
        return True

    def index(self, candidates, identifiers=None):
        prec = self.precision or config.precision
        c = _to_device_matrix(candidates, prec)
        if c.dim() != 2:
            raise ValueError(f"The candidates tensor must be 2D (got {tuple(c.shape)}).")
        if identifiers is not None:
            ident = identifiers if isinstance(identifiers, torch.Tensor) else torch.as_tensor(np.asarray(identifiers))
            if ident.shape[0] != c.shape[0]:
                raise ValueError(f"The candidates and identifiers tensors must have the same number of rows "
                                 f"(got {c.shape[0]} candidates rows and {ident.shape[0]} identifier rows).")
            if ident.dtype != torch.int64:
                ident = ident.to(torch.int64)     # identifiers are the dense int64 item codes
            self._identifiers = ident.to(c.device).contiguous()
        else:
            self._identifiers = None
        self._candidates = c
        self._prec = prec
        return self

    def _embed(self, queries):
        if self.query_model is not None:
            queries = self.query_model(queries)
        return _to_device_matrix(queries, self._prec)

    def __call__(self, queries, k: Optional[int] = None):
        if self._candidates is None:
            raise ValueError("The `index` method must be called first to create the retrieval index.")
        k = int(k if k is not None else self._k)
        q = self._embed(queries)
        return ops.topk_bruteforce(self._prec, q, self._candidates, k, 0, self._identifiers)

    call = __call__

    def query_with_exclusions(self, queries, exclusions, k: Optional[int] = None):
        """Over-fetch k + n_excl, drop excluded identifiers, keep the first k (upstream
        TopK.query_with_exclusions).  The compaction is index bookkeeping done with torch ops."""
        k = int(k if k is not None else self._k)
        excl = exclusions if isinstance(exclusions, torch.Tensor) else torch.as_tensor(np.asarray(exclusions))
        excl = excl.to(self._candidates.device).to(torch.int64)
        n_ex = excl.shape[1]
        kk = min(k + n_ex, self._candidates.shape[0])
        scores, ids = self(queries, k=kk)
        bad = (ids.unsqueeze(2) == excl.unsqueeze(1)).any(dim=2)
        order = torch.argsort(bad.to(torch.int8), dim=1, stable=True)[:, :k]
        return torch.gather(scores, 1, order), torch.gather(ids, 1, order)


class Streaming(BruteForce):
    """tfrs Streaming: same exact result as BruteForce (earlier batch wins ties == lower global
    index); candidates are concatenated at index time and streamed through the same kernel."""

    def __init__(self, query_model=None, k: int = 10, handle_incomplete_batches: bool = True,
                 num_parallel_calls=None, sorted_order: bool = True, name=None, precision=None):
        super().__init__(query_model, k, name, precision)
