"""Candidate-sharded exact retrieval over one 8 x B200 box (SURVEY.md 8e serving row, BASELINE configs[4]: "1M-query x
10M-candidate brute-force top-100 sharded across 8 GPUs vs FAISS flat on host"; the reference's `src/serving` is the
empty package /root/reference/src/serving/__init__.py:1, FAISS is only a dependency pin, /root/reference/pyproject.toml:39).

    index = ShardedBruteForce(k=100, group=dist.group.WORLD).index(my_candidate_shard)
    scores, ids = index(queries)          # queries: the SAME [Q, d] batch on every rank, Q % world == 0
                                          # -> this rank's slice of the answer: rows [rank * Q/world, (rank+1) * Q/world)

Every rank scores all Q queries against its own shard with the brute-force top-k kernel (exact partial lists over
GLOBAL candidate indices), the partial lists of query slice r travel to rank r, and rank r merges its `world` lists by
(score desc, global index asc) -- the same total order as a single-device tf.math.top_k, so the ids are identical to
the 1-GPU result bit for bit.

exchange="peer" (default on NCCL groups): the top-k kernel's re-rank epilogue writes every partial list straight into
the merging rank's symmetric-memory receive area (posted NVLink stores, tt_topk_bruteforce_peer), one flag barrier,
merge -- no collective call and no staging copy.  exchange="collective": torch.distributed all_to_all (NCCL, or gloo in
the CPU tests with an injected `prim`).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from . import ops as _cuda_ops
from .core import config
from .layers.factorized_top_k import TopK, _to_device_matrix


class ShardedBruteForce(TopK):
    def __init__(self, query_model=None, k: int = 10, group=None, precision: Optional[str] = None,
                 exchange: Optional[str] = None, prim=None, name: Optional[str] = None):
        super().__init__(k, name)
        self.query_model = query_model
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.precision = precision
        self.prim = prim or _cuda_ops
        if exchange is None:
            exchange = "peer" if (self.prim is _cuda_ops and dist.get_backend(self.group) == "nccl") else "collective"
        if exchange not in ("peer", "collective"):
            raise ValueError("exchange must be 'peer' or 'collective'")
        self.exchange = exchange
        self._candidates = None
        self._base = 0
        self._identifiers = None
        self._peer = None            # (PeerWorkspace, key, off_s, off_i)

    def is_exact(self) -> bool:
        return True

    # ---- index -------------------------------------------------------------------------------------------------
    def index(self, candidates, identifiers=None, index_base: Optional[int] = None):
        """candidates: THIS rank's shard [n_r, d].  Global candidate index = index_base + local row; by default the
        shards are laid out in rank order (index_base = number of candidates on lower ranks).  identifiers: this
        shard's ids (all-gathered once here so that the merged indices can be translated)."""
        prec = self.precision or config.precision
        if self.prim is _cuda_ops:
            c = _to_device_matrix(candidates, prec)
        else:
            c = candidates if isinstance(candidates, torch.Tensor) else torch.as_tensor(np.asarray(candidates))
        if c.dim() != 2:
            raise ValueError(f"The candidates tensor must be 2D (got {tuple(c.shape)}).")
        counts = [None] * self.world
        dist.all_gather_object(counts, int(c.shape[0]), group=self.group)
        self._counts = counts
        self._base = int(sum(counts[:self.rank])) if index_base is None else int(index_base)
        self._total = int(sum(counts))
        if identifiers is not None:
            ident = identifiers if isinstance(identifiers, torch.Tensor) else torch.as_tensor(np.asarray(identifiers))
            ident = ident.to(device=c.device, dtype=torch.int64).contiguous()
            if ident.shape[0] != c.shape[0]:
                raise ValueError("The candidates and identifiers tensors must have the same number of rows")
            if index_base is not None:
                raise ValueError("identifiers need the default rank-order index layout")
            nmax = max(counts)                                  # shards may be ragged: gather padded, then trim
            padded = torch.zeros((nmax,), dtype=torch.int64, device=c.device)
            padded[:ident.shape[0]] = ident
            parts = [torch.empty((nmax,), dtype=torch.int64, device=c.device) for _ in counts]
            dist.all_gather(parts, padded, group=self.group)
            self._identifiers = torch.cat([p[:n] for p, n in zip(parts, counts)])
        else:
            self._identifiers = None
        self._candidates, self._prec = c, prec
        return self

    # ---- query -------------------------------------------------------------------------------------------------
    def _peer_workspace(self, qpr: int, k: int):
        key = (qpr, k)
        if self._peer is not None and self._peer[1] == key:
            return self._peer
        import torch.distributed._symmetric_memory as symm_mem
        al = lambda n: (n + 1023) // 1024 * 1024
        off_s = _cuda_ops.PeerWorkspace.FLAG_BYTES
        off_i = off_s + al(self.world * qpr * k * 4)
        total = off_i + al(self.world * qpr * k * 8)
        dev = self._candidates.device
        buf = symm_mem.empty((total,), dtype=torch.uint8, device=dev)
        buf.zero_()
        torch.cuda.synchronize()
        handle = symm_mem.rendezvous(buf, self.group)
        bases = torch.tensor([int(p) for p in handle.buffer_ptrs], dtype=torch.int64, device=dev)
        ws = _cuda_ops.PeerWorkspace(buf, bases, self.world, self.rank)
        self._peer = (ws, key, off_s, off_i, handle)
        torch.cuda.synchronize()
        dist.barrier(self.group)
        return self._peer

    def __call__(self, queries, k: Optional[int] = None):
        if self._candidates is None:
            raise ValueError("The `index` method must be called first to create the retrieval index.")
        k = int(k if k is not None else self._k)
        if self.query_model is not None:
            queries = self.query_model(queries)
        prim = self.prim
        q = _to_device_matrix(queries, self._prec) if prim is _cuda_ops else queries
        Q = q.shape[0]
        W = self.world
        if Q % W != 0:
            raise ValueError(f"ShardedBruteForce: the query batch ({Q}) must be a multiple of the group size ({W})")
        if k > self._total:
            raise ValueError(f"k={k} exceeds the number of indexed candidates ({self._total})")
        qpr = Q // W
        k_loc = min(k, int(self._candidates.shape[0]))
        if min(self._counts) < k:
            raise NotImplementedError("ShardedBruteForce: every shard must hold at least k candidates")
        if self.exchange == "peer":
            ws, _key, off_s, off_i, _h = self._peer_workspace(qpr, k)
            # the previous call's merge must have consumed my receive area on every rank before anyone overwrites it
            prim.peer_barrier(ws, 0)
            prim.topk_bruteforce_peer(self._prec, q, self._candidates, k, self._base, ws, qpr, off_s, off_i)
            prim.peer_barrier(ws, 1)
            recv_s = ws.view(off_s, (W, qpr, k), torch.float32)
            recv_i = ws.view(off_i, (W, qpr, k), torch.int64)
        else:
            s, i = prim.topk_bruteforce(self._prec, q, self._candidates, k_loc, self._base, None)
            recv_s, recv_i = torch.empty_like(s), torch.empty_like(i)
            _all_to_all(recv_s, s.contiguous(), self.group)
            _all_to_all(recv_i, i.contiguous(), self.group)
            recv_s, recv_i = recv_s.view(W, qpr, k), recv_i.view(W, qpr, k)
        return prim.topk_merge(recv_s, recv_i, k, 0, self._identifiers)

    call = __call__

    def gather(self, scores, ids):
        """All ranks' slices concatenated in query order (convenience for tests / small batches)."""
        out_s = torch.empty((self.world * scores.shape[0], scores.shape[1]), dtype=scores.dtype, device=scores.device)
        out_i = torch.empty((self.world * ids.shape[0], ids.shape[1]), dtype=ids.dtype, device=ids.device)
        dist.all_gather_into_tensor(out_s, scores.contiguous(), group=self.group)
        dist.all_gather_into_tensor(out_i, ids.contiguous(), group=self.group)
        return out_s, out_i


def _all_to_all(out: torch.Tensor, x: torch.Tensor, group) -> None:
    """x [world * n, ...] -> out[r*n:(r+1)*n] = rank r's x[me*n:(me+1)*n] (gloo has no all_to_all on CPU tensors)."""
    if dist.get_backend(group) == "nccl":
        dist.all_to_all_single(out, x, group=group)
        return
    world, me = dist.get_world_size(group), dist.get_rank(group)
    n = x.shape[0] // world
    bufs = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(bufs, x, group=group)
    for r in range(world):
        out[r * n:(r + 1) * n] = bufs[r][me * n:(me + 1) * n]
