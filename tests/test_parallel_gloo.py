"""world_size-2 and -4 gloo tests (CPU) of the sharded path's routing logic: all-to-all lookups of a
row-sharded table, gradient return to the owners, global-batch negatives.  The local compute
(`prim`) is a numpy/oracle stand-in defined HERE (test infrastructure); the product's default
prim is the CUDA library."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


class CpuPrims:
    """Oracle-backed stand-ins with the signatures of two_tower_b200.ops (fp32 only)."""

    @staticmethod
    def partition_ids(ids, world, capacity=0, overflow_flag=None):
        a = ids.numpy()
        owner = a % world
        order = np.argsort(owner, kind="stable")
        counts = np.bincount(owner, minlength=world)
        starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
        rank_in_bucket = np.arange(len(a)) - starts[owner[order]]
        if capacity:
            pos = owner[order] * capacity + rank_in_bucket
            send = np.full(world * capacity, -1, np.int64)
            fits = rank_in_bucket < capacity
            send[pos[fits]] = (a // world)[order][fits]
            perm = np.full(len(a), -1, np.int64); perm[order[fits]] = pos[fits]
            if overflow_flag is not None and (~fits).any():
                overflow_flag[0] = 1
        else:
            send = (a // world)[order]
            perm = np.empty(len(a), np.int64); perm[order] = np.arange(len(a))
        return torch.from_numpy(send), torch.from_numpy(perm), torch.from_numpy(counts)

    @staticmethod
    def embedding_gather(table, ids, out_dtype=torch.float32):
        out = torch.zeros((ids.numel(), table.shape[1]), dtype=out_dtype)
        ok = ids >= 0
        out[ok] = table[ids[ok]].to(out_dtype)
        return out

    @staticmethod
    def permute_rows(x, perm, inverse, out_rows=0, zero_fill=False):
        n = perm.numel()
        ok = perm >= 0
        if inverse:
            out = torch.zeros((n, x.shape[1]), dtype=x.dtype)
            out[ok] = x[perm[ok]]
        else:
            out = torch.zeros((out_rows or n, x.shape[1]), dtype=x.dtype)
            out[perm[ok]] = x[:n][ok]
        return out

    @staticmethod
    def _rot(nq, nc, off):
        return np.concatenate([np.arange(off, off + nq), np.arange(0, off), np.arange(off + nq, nc)])

    @classmethod
    def retrieval_loss_fwd(cls, prec, q, c, inv_t, label_offset=0, w=None, logq=None, ids=None):
        import oracle
        perm = cls._rot(q.shape[0], c.shape[0], label_offset)
        r = oracle.retrieval_loss_and_grads(q.numpy().astype(np.float64), c.numpy().astype(np.float64)[perm], temperature=1.0 / inv_t)
        return torch.tensor([r["loss"]], dtype=torch.float64), torch.from_numpy(r["lse"]), torch.from_numpy(r["pos"])

    @classmethod
    def retrieval_loss_bwd(cls, prec, q, c, inv_t, lse, label_offset=0, w=None, logq=None, ids=None,
                           grad_scale=1.0, want_bf16=(False, False)):
        import oracle
        perm = cls._rot(q.shape[0], c.shape[0], label_offset)
        r = oracle.retrieval_loss_and_grads(q.numpy().astype(np.float64), c.numpy().astype(np.float64)[perm], temperature=1.0 / inv_t)
        dc = np.empty_like(r["dc"]); dc[perm] = r["dc"]
        return dict(dq=torch.from_numpy(r["dq"]), dc=torch.from_numpy(dc), dq_bf16=None, dc_bf16=None)


    @staticmethod
    def topk_bruteforce(prec, q, c, k, cand_index_base=0, identifiers=None, uncertain=None):
        import oracle
        s, i = oracle.brute_force_topk(q.numpy(), c.numpy(), k, score_dtype=np.float32)
        return torch.from_numpy(s), torch.from_numpy(i + cand_index_base)

    @staticmethod
    def topk_merge(scores, ids, k_out, index_base=0, identifiers=None):
        import oracle
        L = scores.shape[0]
        s, i = oracle.topk_merge([scores[l].numpy() for l in range(L)], [ids[l].numpy() for l in range(L)], k_out)
        i = i + index_base if identifiers is None else identifiers.numpy()[i]
        return torch.from_numpy(s), torch.from_numpy(i)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, WORLD, port, results):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        import oracle
        import two_tower_b200 as tt
        from two_tower_b200 import parallel, synth
        from two_tower_b200.core import GradientTape, Tensor
        tt.set_precision("fp32")
        coll = parallel.Collectives(None)
        V, d, b = 1003, 16, 96
        rng_all = synth.rng_for(11)
        table = oracle.keras_uniform(rng_all, (V, d)).astype(np.float64)
        ids_all = [synth.draw_ids(synth.rng_for(100 + r), b, V, zipf=1.2) for r in range(WORLD)]
        grads_all = [synth.rng_for(200 + r).normal(size=(b, d)) for r in range(WORLD)]
        shard = torch.from_numpy(table[rank::WORLD].copy())

        # collectives emulation
        x = torch.arange(WORLD * 3, dtype=torch.float64).reshape(WORLD * 3, 1) + 100 * rank
        y = coll.all_to_all(x)
        for r in range(WORLD):
            assert torch.equal(y[r * 3:(r + 1) * 3], torch.arange(rank * 3, rank * 3 + 3, dtype=torch.float64).reshape(3, 1) + 100 * r)

        for cap_factor in (None, 1.6):
            cap = parallel.bucket_capacity(b, WORLD, cap_factor)
            flag = torch.zeros(1, dtype=torch.int32)
            rows, ctx = parallel.exchange_lookup(CpuPrims, coll, shard, torch.from_numpy(ids_all[rank]), cap, torch.float64, flag)
            assert int(flag.item()) == 0
            assert np.array_equal(rows.numpy(), table[ids_all[rank]])             # lookups exact
            # per-owner bucket = that owner's ids from every rank, each rank's in batch order (stable)
            recv_ids = ctx[0].numpy().reshape(WORLD, cap)
            for r in range(WORLD):
                mine = ids_all[r][ids_all[r] % WORLD == rank] // WORLD
                assert np.array_equal(recv_ids[r][:len(mine)], mine) and (recv_ids[r][len(mine):] == -1).all()
            local_ids, g = parallel.exchange_grads(CpuPrims, coll, ctx, torch.from_numpy(grads_all[rank]), cap)
            ok = local_ids.numpy() >= 0
            acc0 = np.full(shard.shape, 0.1)
            t_new, a_new, uniq = oracle.adagrad_sparse(shard.numpy(), acc0, local_ids.numpy()[ok], g.numpy()[ok], lr=0.1)
            t_ref, a_ref, uniq_ref = oracle.adagrad_sparse(table, np.full(table.shape, 0.1), np.concatenate(ids_all),
                                                           np.concatenate(grads_all), lr=0.1)
            np.testing.assert_allclose(t_new, t_ref[rank::WORLD], rtol=1e-12, atol=1e-15)
            # gradient row set of the shard is exactly the owned part of the global set
            assert np.array_equal(np.sort(uniq * WORLD + rank), np.sort(uniq_ref[uniq_ref % WORLD == rank]))

        # overflow is reported, not silently dropped
        flag = torch.zeros(1, dtype=torch.int32)
        parallel.exchange_lookup(CpuPrims, coll, shard, torch.from_numpy(ids_all[rank]), 4, torch.float64, flag)
        assert int(flag.item()) == 1

        # global-batch negatives
        q_all = [synth.rng_for(300 + r).normal(size=(b, d)) * 0.3 for r in range(WORLD)]
        c_all = [synth.rng_for(400 + r).normal(size=(b, d)) * 0.3 for r in range(WORLD)]
        task = tt.tasks.Retrieval(temperature=0.5, process_group=dist.group.WORLD)
        q = Tensor(f32=torch.from_numpy(q_all[rank])); c = Tensor(f32=torch.from_numpy(c_all[rank]))
        with GradientTape() as tape:
            loss = parallel.global_retrieval(task, q, c, 2.0, None, None, None, prim=CpuPrims)
            tape.gradient(loss, [])
        total = loss.value.clone(); dist.all_reduce(total)
        ref = oracle.retrieval_loss_and_grads(np.concatenate(q_all), np.concatenate(c_all), temperature=0.5)
        assert float(total.item()) == pytest.approx(ref["loss"], rel=1e-12)
        np.testing.assert_allclose(q.grad["f32"].numpy(), ref["dq"][rank * b:(rank + 1) * b], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(c.grad["f32"].numpy(), ref["dc"][rank * b:(rank + 1) * b], rtol=1e-10, atol=1e-13)
        results[rank] = "ok"
    except Exception as e:  # surfaced by the parent
        import traceback
        results[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("WORLD", [2, 4])
def test_sharded_lookup_gradients_and_global_negatives(WORLD):
    mgr = mp.Manager()
    results = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(WORLD, port, results), nprocs=WORLD, join=True)
    for r in range(WORLD):
        assert results.get(r) == "ok", results.get(r)


def _serving_worker(rank, WORLD, port, results):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        import oracle
        import two_tower_b200 as tt
        from two_tower_b200 import synth
        rng = synth.rng_for(91)
        d, k, nq = 16, 10, 6 * WORLD
        sizes = [300, 211, 157, 243][:WORLD]                 # ragged shards
        cand = synth.exact_matrix(rng, sum(sizes), d, 2)     # dyadic: real ties, also ACROSS shards
        q = synth.exact_matrix(rng, nq, d, 2)
        ident = rng.permutation(sum(sizes)).astype(np.int64) + 1000
        lo = sum(sizes[:rank])
        ref_s, ref_i = oracle.brute_force_topk(q, cand, k, score_dtype=np.float32)
        for use_ident in (False, True):
            index = tt.serving.ShardedBruteForce(k=k, group=dist.group.WORLD, precision="fp32", prim=CpuPrims)
            assert index.exchange == "collective"
            index.index(torch.from_numpy(cand[lo:lo + sizes[rank]]),
                        identifiers=torch.from_numpy(ident[lo:lo + sizes[rank]]) if use_ident else None)
            s, i = index(torch.from_numpy(q))
            a, b = rank * (nq // WORLD), (rank + 1) * (nq // WORLD)
            want = ident[ref_i[a:b]] if use_ident else ref_i[a:b]
            assert np.array_equal(i.numpy(), want)
            assert np.array_equal(s.numpy(), ref_s[a:b])
        with pytest.raises(ValueError, match="multiple of the group size"):
            index(torch.from_numpy(q[:nq - 1]))
        results[rank] = "ok"
    except Exception:
        import traceback
        results[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("WORLD", [2, 4])
def test_candidate_sharded_topk_routing(WORLD):
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_serving_worker, args=(WORLD, _free_port(), results), nprocs=WORLD, join=True)
    for r in range(WORLD):
        assert results.get(r) == "ok", results.get(r)
