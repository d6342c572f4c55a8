"""tfrs.tasks.Retrieval: in-batch softmax retrieval loss, fused forward and
backward kernels, logits never in HBM."""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np
import torch

from . import ops
from .core import GradientTape, Scalar, Tensor, config


def _as_f32_device(x, dev):
    if x is None:
        return None
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
    return t.to(device=dev, dtype=torch.float32).contiguous()


class Retrieval:
    """tfrs.tasks.Retrieval(loss=None, metrics=None, batch_metrics=None, loss_metrics=None,
    temperature=None, num_hard_negatives=None, remove_accidental_hits=False).

    Default (and only hot-path) loss: CategoricalCrossentropy(from_logits=True, reduction=SUM)
    with labels = eye(num_queries, num_candidates).  ``process_group`` (extension, SURVEY.md
    8e): all-gather the candidate embeddings over the group so negatives span the global
    batch; the result equals the single-device loss on the concatenated batch.
    """

    def __init__(self, loss=None, metrics=None, batch_metrics=None, loss_metrics=None,
                 temperature: Optional[float] = None, num_hard_negatives: Optional[int] = None,
                 remove_accidental_hits: bool = False, name: Optional[str] = None, process_group=None):
        if loss is not None:
            raise NotImplementedError("only the default CategoricalCrossentropy(from_logits=True, reduction=SUM) is built")
        if num_hard_negatives is not None:
            if int(num_hard_negatives) < 1:
                raise ValueError("num_hard_negatives must be a positive integer")
            if process_group is not None:
                raise NotImplementedError("num_hard_negatives with global-batch negatives (process_group) is not built")
        self._num_hard_negatives = None if num_hard_negatives is None else int(num_hard_negatives)
        self._factorized_metrics = metrics
        self._batch_metrics = batch_metrics
        self._loss_metrics = loss_metrics
        self._temperature = temperature
        self._remove_accidental_hits = bool(remove_accidental_hits)
        self.name = name or "retrieval"
        self.process_group = process_group

    @property
    def factorized_metrics(self):
        return self._factorized_metrics

    @factorized_metrics.setter
    def factorized_metrics(self, value):
        self._factorized_metrics = value

    def __call__(self, query_embeddings: Tensor, candidate_embeddings: Tensor, sample_weight=None,
                 candidate_sampling_probability=None, candidate_ids=None, compute_metrics: bool = True,
                 compute_batch_metrics: bool = True) -> Scalar:
        q, c = query_embeddings, candidate_embeddings
        prec = config.precision
        dev = (q.f32 if q.f32 is not None else q.bf16).device
        nq = q.shape[0]
        inv_t = 1.0 if self._temperature is None else 1.0 / float(self._temperature)
        w = _as_f32_device(sample_weight, dev)
        logq = None
        if candidate_sampling_probability is not None:
            p = _as_f32_device(candidate_sampling_probability, dev)
            logq = torch.log(torch.clamp(p, 1e-6, 1.0))      # [nc] input preparation
        ids = None
        if self._remove_accidental_hits:
            if candidate_ids is None:
                raise ValueError("When accidental hit removal is enabled, candidate ids must be supplied.")
            ids = candidate_ids if isinstance(candidate_ids, torch.Tensor) else torch.as_tensor(np.asarray(candidate_ids))
            ids = ids.to(device=dev, dtype=torch.int64).contiguous()

        if self.process_group is not None:
            from . import parallel
            return parallel.global_retrieval(self, q, c, inv_t, w, logq, ids)

        if self._num_hard_negatives is not None:
            # tfrs HardNegativeMining: the positive + the n highest-scoring negatives of each row.  Selection by the
            # brute-force top-k kernel (k = n + 1), loss and gradients on the gathered logits (csrc/hard_negatives.cu).
            if logq is not None or ids is not None:
                raise NotImplementedError("num_hard_negatives together with candidate_sampling_probability / "
                                          "remove_accidental_hits is not built (the selection runs on the plain scores)")
            qm, cm = (q.f32, c.f32) if prec == "fp32" else (q.bf16, c.bf16)
            sel = ops.select_hard_negatives(prec, qm, cm, self._num_hard_negatives, 0)
            loss, lse, _pos, scores = ops.hard_negative_loss_fwd(prec, qm, cm, sel, inv_t, w)

            def backward_hard():
                dq, dc = ops.hard_negative_loss_bwd(prec, qm, cm, sel, inv_t, scores, lse, w, 1.0)
                for t, g in ((q, dq), (c, dc)):
                    if "parts" in t.grad_formats:
                        t.grad = dict(parts=g.reshape(1, *g.shape))
                    else:
                        t.grad = dict(f32=g, bf16=ops.cast_f32_to_bf16(g) if (prec == "bf16" and "bf16" in t.grad_formats) else None)

            GradientTape.record(backward_hard)
            if compute_metrics and self._factorized_metrics is not None:
                self._factorized_metrics.update_state(q, _first_rows(c, nq),
                                                      true_candidate_ids=None if candidate_ids is None else candidate_ids,
                                                      sample_weight=sample_weight)
            return Scalar(loss)

        if prec == "fp32":
            qm, cm = q.f32, c.f32
        else:
            qm, cm = q.bf16, c.bf16
        # bf16 training step: the forward also accumulates dQ (one exponential and one S GEMM fewer per logit)
        fused = (prec == "bf16" and GradientTape.current() is not None and logq is None and ids is None
                 and ops.retrieval_fwd_dq_supported(nq, cm.shape[0], qm.shape[1]))
        if fused:
            loss, lse, _pos, dq_fused, _fwd_ws = ops.retrieval_loss_fwd_dq(qm, cm, inv_t, 0, w, fork=True)
        else:
            loss, lse, _pos = ops.retrieval_loss_fwd(prec, qm, cm, inv_t, 0, w, logq, ids)

        def backward():
            bf = prec == "bf16"
            if fused:
                # Two alternatives were built and measured SLOWER on B200 (cfg2 step, 163 us with this form):
                # forking the fold onto a side stream next to the dC pass (187 us: its blocks take issue slots from the
                # softmax warps) and feeding the dC pass -lse2 by broadcast global loads instead of the shared-memory
                # staging (retrieval_loss_bwd_dc_fused, 173 us).
                _none, dc_parts = ops.retrieval_loss_bwd_parts(qm, cm, inv_t, lse, 0, w, None, None, 1.0, want_dq=False)
                ops.join_side_work()                 # the forked loss summation is done before anything downstream
                if "parts" in q.grad_formats:
                    q.grad = dict(parts=dq_fused.reshape(1, *dq_fused.shape))
                else:
                    q.grad = dict(f32=dq_fused, bf16=ops.cast_f32_to_bf16(dq_fused) if "bf16" in q.grad_formats else None)
                if "parts" in c.grad_formats:
                    c.grad = dict(parts=dc_parts)
                else:
                    f, b = ops.combine_parts(dc_parts, True, "bf16" in c.grad_formats)
                    c.grad = dict(f32=f, bf16=b)
                return
            if bf and ("parts" in q.grad_formats or "parts" in c.grad_formats):
                # the consumer (fused tower backward) folds the split partials itself
                dq_parts, dc_parts = ops.retrieval_loss_bwd_parts(qm, cm, inv_t, lse, 0, w, logq, ids, 1.0)
                for t, parts in ((q, dq_parts), (c, dc_parts)):
                    if "parts" in t.grad_formats:
                        t.grad = dict(parts=parts)
                    else:
                        f, b = ops.combine_parts(parts, True, "bf16" in t.grad_formats)
                        t.grad = dict(f32=f, bf16=b)
                return
            r = ops.retrieval_loss_bwd(prec, qm, cm, inv_t, lse, 0, w, logq, ids, 1.0,
                                       want_bf16=(bf and "bf16" in q.grad_formats, bf and "bf16" in c.grad_formats))
            q.grad = dict(f32=r["dq"], bf16=r["dq_bf16"])
            c.grad = dict(f32=r["dc"], bf16=r["dc_bf16"])

        GradientTape.record(backward)

        if compute_metrics and self._factorized_metrics is not None:
            self._factorized_metrics.update_state(q, _first_rows(c, nq),
                                                  true_candidate_ids=None if candidate_ids is None else candidate_ids,
                                                  sample_weight=sample_weight)
        return Scalar(loss)

    call = __call__


def _first_rows(c: Tensor, n: int) -> Tensor:
    if c.shape[0] == n:
        return c
    return Tensor(f32=None if c.f32 is None else c.f32[:n].contiguous(),
                  bf16=None if c.bf16 is None else c.bf16[:n].contiguous())
