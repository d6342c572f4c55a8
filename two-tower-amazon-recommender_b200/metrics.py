"""tfrs.metrics.FactorizedTopK: top-k categorical accuracy of the true
candidate against the whole corpus, one weighted mean per k
(ks: /root/reference/configs/data_config.yaml:71)."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import ops
from .core import Tensor, config
from .layers.factorized_top_k import BruteForce, Streaming, TopK, _to_device_matrix


class FactorizedTopK:
    def __init__(self, candidates, ks: Sequence[int] = (1, 5, 10, 50, 100), name: str = "factorized_top_k",
                 precision: Optional[str] = None):
        self.name = name
        self._ks = tuple(int(k) for k in ks)
        if not 1 <= len(self._ks) <= 8:
            raise ValueError("FactorizedTopK supports 1..8 values of k")
        if isinstance(candidates, TopK):
            self._candidates = candidates
        else:
            # a dataset / array of candidate embeddings is wrapped in Streaming(k=max(ks))
            idx = Streaming(k=max(self._ks), precision=precision)
            if isinstance(candidates, (torch.Tensor, np.ndarray, Tensor)):
                idx.index(candidates)
            else:
                idx.index_from_dataset(candidates)
            self._candidates = idx
        self._hits = None
        self._weight = None

    @property
    def ks(self):
        return self._ks

    def reset_states(self):
        if self._hits is not None:
            self._hits.zero_()
            self._weight.zero_()

    reset_state = reset_states

    def update_state(self, query_embeddings, true_candidate_embeddings, true_candidate_ids=None,
                     sample_weight=None):
        prec = self._candidates._prec
        q = _to_device_matrix(query_embeddings, prec)
        dev = q.device
        if self._hits is None:
            self._hits = torch.zeros(8, dtype=torch.float32, device=dev)
            self._weight = torch.zeros(1, dtype=torch.float32, device=dev)
        kmax = max(self._ks)
        scores, ids = self._candidates(q, k=kmax)
        w = None
        if sample_weight is not None:
            w = (sample_weight if isinstance(sample_weight, torch.Tensor) else torch.as_tensor(np.asarray(sample_weight)))
            w = w.to(device=dev, dtype=torch.float32).contiguous()
        if true_candidate_ids is None:
            c = _to_device_matrix(true_candidate_embeddings, prec)
            positive = ops.rowwise_dot(prec, q, c)
            ops.topk_hits(positive, scores, None, None, w, self._ks, self._hits, self._weight)
        else:
            t = true_candidate_ids if isinstance(true_candidate_ids, torch.Tensor) else torch.as_tensor(np.asarray(true_candidate_ids))
            t = t.to(device=dev, dtype=torch.int64).contiguous().reshape(-1)
            ops.topk_hits(None, None, ids, t, w, self._ks, self._hits, self._weight)

    def result(self):
        if self._hits is None:
            return {f"{self.name}/top_{k}_categorical_accuracy": 0.0 for k in self._ks}
        h = self._hits.cpu().numpy()
        wsum = float(self._weight.cpu().numpy()[0])
        return {f"{self.name}/top_{k}_categorical_accuracy": (float(h[i]) / wsum if wsum else 0.0)
                for i, k in enumerate(self._ks)}
