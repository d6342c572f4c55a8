"""Evaluation loop: corpus-wide retrieval metrics of a two-tower model every
`validation_freq` epochs and early stopping with `patience`, parameterised by the `model:` block of the reference's
`configs/data_config.yaml` (`training.patience: 5`, `training.validation_freq: 1`, `retrieval.top_k_eval:
[1, 5, 10, 20, 50, 100]`, lines 61-71) -- the reference's `src/evaluation` package is empty.

The scoring itself is the brute-force top-k kernel (`layers.factorized_top_k.BruteForce`, tt_topk_bruteforce): every
validation query is scored against ALL item embeddings; what happens here is bookkeeping on the [Q, max k] id lists.
With one relevant item per query: Recall@k = hit rate = tfrs' `factorized_top_k/top_k_categorical_accuracy`;
NDCG@k = 1 / log2(rank + 2) for a hit at 0-based rank < k, else 0; MRR over the top max(k)."""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Sequence

import numpy as np
import torch

from .layers.factorized_top_k import BruteForce

TOP_K_EVAL = (1, 5, 10, 20, 50, 100)          # configs/data_config.yaml:71


def ranks_of_true_ids(topk_ids: torch.Tensor, true_ids: torch.Tensor) -> torch.Tensor:
    """0-based rank of true_ids[i] inside topk_ids[i, :], or K if absent.  [Q, K], [Q] -> [Q] int64."""
    match = topk_ids == true_ids.reshape(-1, 1)
    K = topk_ids.shape[1]
    first = torch.where(match.any(dim=1), match.to(torch.int64).argmax(dim=1), torch.full_like(true_ids.reshape(-1), K))
    return first


def metrics_from_ranks(ranks: torch.Tensor, K: int, ks: Sequence[int]) -> Dict[str, float]:
    r = ranks.to(torch.float64)
    out = {}
    for k in ks:
        hit = ranks < min(k, K)
        out[f"recall@{k}"] = float(hit.to(torch.float64).mean().item())
        out[f"ndcg@{k}"] = float(torch.where(hit, 1.0 / torch.log2(r + 2.0), torch.zeros_like(r)).mean().item())
        out[f"factorized_top_k/top_{k}_categorical_accuracy"] = out[f"recall@{k}"]
    found = ranks < K
    out["mrr"] = float(torch.where(found, 1.0 / (r + 1.0), torch.zeros_like(r)).mean().item())
    return out


class RetrievalEvaluator:
    """Recall@k / NDCG@k / MRR of `user_model` queries against the whole item corpus embedded by `item_model`.

        ev = RetrievalEvaluator(model.user_model, model.item_model, num_items=ds.num_items)
        metrics = ev.evaluate(validation_batches)          # dict: recall@1 ... ndcg@100, mrr, n_queries
    """

    def __init__(self, user_model, item_model, num_items: int, ks: Sequence[int] = TOP_K_EVAL,
                 query_key: str = "user_id_encoded", item_key: str = "item_id_encoded", corpus_batch: int = 65536):
        self.user_model, self.item_model = user_model, item_model
        self.num_items, self.ks = int(num_items), tuple(int(k) for k in ks)
        self.query_key, self.item_key, self.corpus_batch = query_key, item_key, int(corpus_batch)

    def build_index(self) -> BruteForce:
        dev = torch.device("cuda", torch.cuda.current_device())
        chunks = []
        for lo in range(0, self.num_items, self.corpus_batch):
            ids = torch.arange(lo, min(lo + self.corpus_batch, self.num_items), dtype=torch.int64, device=dev)
            e = self.item_model(ids)
            chunks.append((e.bf16 if e.bf16 is not None else e.f32).clone())
        emb = torch.cat(chunks)
        return BruteForce(k=min(max(self.ks), self.num_items)).index(emb)          # identifiers = row index = item id

    def evaluate(self, batches: Iterable[dict], index: Optional[BruteForce] = None) -> Dict[str, float]:
        index = index or self.build_index()
        K = min(max(self.ks), self.num_items)
        ranks = []
        for batch in batches:
            q = self.user_model(batch[self.query_key])
            _scores, ids = index(q, k=K)
            true = batch[self.item_key]
            true = (true if isinstance(true, torch.Tensor) else torch.as_tensor(np.asarray(true))).to(ids.device, torch.int64)
            ranks.append(ranks_of_true_ids(ids, true))
        if not ranks:
            return {}
        ranks = torch.cat(ranks)
        out = metrics_from_ranks(ranks, K, self.ks)
        out["n_queries"] = int(ranks.numel())
        return out


class EarlyStopping:
    """Stop when `monitor` has not improved for `patience` validation rounds (data_config.yaml:64: patience 5)."""

    def __init__(self, monitor: str = "recall@10", patience: int = 5, mode: str = "max", min_delta: float = 0.0):
        if mode not in ("max", "min"):
            raise ValueError("mode must be 'max' or 'min'")
        self.monitor, self.patience, self.mode, self.min_delta = monitor, int(patience), mode, float(min_delta)
        self.best, self.best_epoch, self.wait = None, -1, 0

    def update(self, epoch: int, metrics: Dict[str, float]) -> bool:
        """Record a validation result; returns True when training should stop."""
        if self.monitor not in metrics:
            raise KeyError(f"EarlyStopping: metric {self.monitor!r} not in {sorted(metrics)}")
        v = metrics[self.monitor]
        better = self.best is None or (v > self.best + self.min_delta if self.mode == "max" else v < self.best - self.min_delta)
        if better:
            self.best, self.best_epoch, self.wait = v, epoch, 0
            return False
        self.wait += 1
        return self.wait >= self.patience


def fit(model, train_batches, epochs: int = 50, validation_batches=None, evaluator: Optional[RetrievalEvaluator] = None,
        validation_freq: int = 1, early_stopping: Optional[EarlyStopping] = None, step=None) -> Dict[str, list]:
    """The training loop the reference's config describes (epochs 50, validation_freq 1, patience 5): `step` defaults
    to `model.train_step` (pass a `GraphedStep` for CUDA-graph replay); validation every `validation_freq` epochs
    through `evaluator`; stops early when `early_stopping` says so.  Returns the history."""
    step = step or model.train_step
    hist = {"loss": [], "val": [], "val_epoch": [], "stopped_epoch": None}
    for epoch in range(epochs):
        total, n = None, 0
        for batch in train_batches:
            loss = step(batch)["loss"]
            total = loss.clone() if total is None else total + loss       # device-side sum: no sync per step
            n += 1
        hist["loss"].append(float(total.item()) / max(n, 1) if total is not None else float("nan"))
        if evaluator is not None and validation_batches is not None and (epoch + 1) % max(1, validation_freq) == 0:
            m = evaluator.evaluate(validation_batches)
            hist["val"].append(m)
            hist["val_epoch"].append(epoch)
            if early_stopping is not None and early_stopping.update(epoch, m):
                hist["stopped_epoch"] = epoch
                break
    return hist
