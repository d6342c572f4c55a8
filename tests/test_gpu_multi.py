"""NCCL tests of the sharded paths at world = min(visible GPUs, 4): row-sharded tables with P2P / all-to-all
lookups and gradients + all-gathered candidates must reproduce the oracle's step on the concatenated global batch,
and candidate-sharded serving must return the ids of the single-device search.  On a 1-GPU box the same code runs
at world = 1 (symmetric-memory workspace, flag barriers, owner-filtered optimizer launch, peer-memory kernels all
execute; only the NVLink hop is missing); TT_TEST_WORLD overrides the world size."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _world():
    n = torch.cuda.device_count()
    w = int(os.environ.get("TT_TEST_WORLD", "0")) or min(n, 4)
    return max(1, min(w, n))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, WORLD, port, precision, peer, dim, mlp, results):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=torch.device("cuda", rank))
    try:
        import oracle
        import two_tower_b200 as tt
        from two_tower_b200 import parallel, synth
        tt.set_precision(precision)
        tol = 1e-5 if precision == "fp32" else 2e-2
        cfg = synth.Config("tiny", 77, 256, dim, 3001, 2003, mlp, 0.5)
        model = parallel.build_sharded_two_tower(cfg, dist.group.WORLD, lr=0.05, capacity_factor=None, peer=peer)
        assert type(model.user_model.layers[0]).__name__ == ("PeerShardedEmbedding" if peer else "ShardedEmbedding")
        assert (model.exchange is not None) == (peer == "exchange")
        batches = [synth.make_batch(cfg, 10 + r) for r in range(WORLD)]
        model.test_step(batches[rank])                                  # builds the Dense layers
        # one global set of weights: full tables from a common seed, Dense weights from rank 0
        rng = synth.rng_for(5)
        U = oracle.keras_uniform(rng, (cfg.v_user, cfg.dim)); I = oracle.keras_uniform(rng, (cfg.v_item, cfg.dim))
        model.user_model.layers[0].load_full_table(U); model.item_model.layers[0].load_full_table(I)
        qs = oracle.TowerSpec([("user_id_encoded", "id", cfg.v_user, None)], cfg.dim, cfg.mlp)
        cs = oracle.TowerSpec([("item_id_encoded", "id", cfg.v_item, None)], cfg.dim, cfg.mlp)
        def dense_params(seq):
            ks, bs = [], []
            for l in seq.layers[1:]:
                for v in (l.kernel, l.bias):
                    dist.broadcast(v.value, src=0)
                    v.refresh_shadows()
                ks.append(l.kernel.numpy().astype(np.float64)); bs.append(l.bias.numpy().astype(np.float64))
            return ks, bs
        qk, qb = dense_params(model.user_model); ck, cb = dense_params(model.item_model)
        qp = {"tables": {"user_id_encoded": U.astype(np.float64)}, "kernels": qk, "biases": qb}
        cp = {"tables": {"item_id_encoded": I.astype(np.float64)}, "kernels": ck, "biases": cb}
        mk = lambda p: {"tables": {k: np.full(v.shape, 0.1) for k, v in p["tables"].items()},
                        "kernels": [np.full(k.shape, 0.1) for k in p["kernels"]], "biases": [np.full(x.shape, 0.1) for x in p["biases"]]}
        out = model.train_step(batches[rank])
        gq = {"user_id_encoded": np.concatenate([b["user_id_encoded"] for b in batches])}
        gc = {"item_id_encoded": np.concatenate([b["item_id_encoded"] for b in batches])}
        ref = oracle.two_tower_train_step(qs, cs, qp, cp, mk(qp), mk(cp), gq, gc, temperature=cfg.temperature, lr=0.05,
                                          bf16=precision == "bf16")
        assert abs(float(out["loss"].item()) - ref["loss"]) <= tol * abs(ref["loss"]), (float(out["loss"].item()), ref["loss"])
        shard = model.user_model.layers[0].embeddings.numpy()
        want = qp["tables"]["user_id_encoded"][rank::WORLD]
        touched = np.unique(gq["user_id_encoded"]); mine = touched[touched % WORLD == rank] // WORLD
        rest = np.setdiff1d(np.arange(want.shape[0]), mine)
        assert np.array_equal(shard[rest], U[rank::WORLD][rest])                      # gradient row set exact per shard
        assert (np.abs(shard[mine] - U[rank::WORLD][mine]).max(axis=1) > 0).all()
        if precision == "fp32":
            err = np.abs(shard[:want.shape[0]] - want).max() / np.abs(want).max()
            assert err < 10 * tol, err
            k0 = model.user_model.layers[1].kernel.numpy()
            assert np.abs(k0 - qp["kernels"][0]).max() / np.abs(qp["kernels"][0]).max() < 10 * tol
        if peer == "exchange":
            # the peer-memory exchange against the NCCL collectives over several steps (barrier epochs, workspace
            # reuse): same weights, same batches -> same losses and the same table shard
            ref_model = parallel.build_sharded_two_tower(cfg, dist.group.WORLD, lr=0.05, capacity_factor=None, peer=True)
            ref_model.test_step(batches[rank])
            for m in (model, ref_model):
                m.user_model.layers[0].load_full_table(U); m.item_model.layers[0].load_full_table(I)
                for seq in (m.user_model, m.item_model):
                    seq.layers[0].embeddings.slots.clear()
            for seq_a, seq_b in ((model.user_model, ref_model.user_model), (model.item_model, ref_model.item_model)):
                for la, lb in zip(seq_a.layers[1:], seq_b.layers[1:]):
                    for va, vb in ((la.kernel, lb.kernel), (la.bias, lb.bias)):
                        vb.value.copy_(va.value); vb.refresh_shadows()
                        for k, sl in va.slots.items():
                            if isinstance(sl, torch.Tensor):
                                vb.slots[k] = sl.clone()
            t0 = model.user_model.layers[0].embeddings.numpy().copy()
            for s in range(4):
                bt = synth.make_batch(cfg, 100 + 10 * s + rank)
                la, lb = float(model.train_step(bt)["loss"].item()), float(ref_model.train_step(bt)["loss"].item())
                assert abs(la - lb) <= 1e-4 * abs(lb), (s, la, lb)
                if s == 0:
                    # Same weights in -> the same shard out, up to the summation order of the exchanges (fp32 atomics over
                    # duplicate ids, slot order of the dC / dense sums): stated relative to the UPDATE.  Only the first
                    # step is compared: this random-init model scores every candidate alike (loss = B ln B_glob), so its
                    # gradients are differences of nearly equal bf16 vectors and the two equivalent variants drift apart
                    # by a whole update within three more steps at world 4 (measured; world 2 happens to stay within 1e-5).
                    ta, tb = model.user_model.layers[0].embeddings.numpy(), ref_model.user_model.layers[0].embeddings.numpy()
                    upd = np.abs(tb - t0).max()
                    assert upd > 0 and np.abs(ta - tb).max() <= 2e-2 * upd, (np.abs(ta - tb).max(), upd)
            g = model.make_graphed_train_step({k: torch.from_numpy(v).cuda() for k, v in batches[rank].items()}, warmup=2)
            for s in range(3):                                           # replay: epochs live in device memory
                out = g({k: torch.from_numpy(v).cuda() for k, v in synth.make_batch(cfg, 200 + 10 * s + rank).items()})
            assert np.isfinite(float(out["loss"].item()))
        results[rank] = "ok"
    except Exception:
        import traceback
        results[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


# peer=False: NCCL all-to-all lookups; peer=True: shards in symmetric memory, P2P gather inside the kernels.
# (128, (256, 128)) in bf16 is the shape the fused tower kernels take.
@pytest.mark.parametrize("precision,peer,dim,mlp", [("fp32", False, 64, (128, 64)), ("bf16", False, 64, (128, 64)),
                                                     ("fp32", True, 64, (128, 64)), ("bf16", True, 128, (256, 128)),
                                                     ("bf16", False, 128, (256, 128)), ("bf16", "exchange", 128, (256, 128))])
def test_sharded_two_tower_matches_global_batch_oracle(precision, peer, dim, mlp):
    WORLD = _world()
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(WORLD, _free_port(), precision, peer, dim, mlp, results), nprocs=WORLD, join=True)
    for r in range(WORLD):
        assert results.get(r) == "ok", results.get(r)


def _worker_serving(rank, WORLD, port, results):
    """Candidate-sharded top-k (BASELINE configs[4]): both exchanges against the single-device kernel and the oracle."""
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=torch.device("cuda", rank))
    try:
        import oracle
        import two_tower_b200 as tt
        from two_tower_b200 import synth
        d, k = 128, 100
        nq, per = 64 * WORLD, 20000 + 512
        rng = synth.rng_for(55)
        cand = oracle.bf16_round((rng.normal(size=(per * WORLD, d)) / np.sqrt(d)).astype(np.float32))
        dup = per if WORLD > 1 else 100
        cand[dup:dup + 3] = cand[0:3]                       # exact duplicates (across shards): ties -> lower GLOBAL index
        q = oracle.bf16_round((rng.normal(size=(nq, d)) / np.sqrt(d)).astype(np.float32))
        ident = (np.arange(per * WORLD, dtype=np.int64) * 7 + 3)
        ref_s, ref_i = oracle.brute_force_topk(q, cand, k, score_dtype=np.float32, block=16384)
        mine = torch.as_tensor(cand[rank * per:(rank + 1) * per]).cuda().to(torch.bfloat16)
        qd = torch.as_tensor(q).cuda().to(torch.bfloat16)
        lo, hi = rank * (nq // WORLD), (rank + 1) * (nq // WORLD)
        for exchange in ("peer", "collective"):
            for use_ident in (False, True):
                index = tt.serving.ShardedBruteForce(k=k, group=dist.group.WORLD, precision="bf16", exchange=exchange)
                index.index(mine, identifiers=torch.as_tensor(ident[rank * per:(rank + 1) * per]) if use_ident else None)
                for rep in range(3):                        # repeated calls re-use the receive area (barrier epochs)
                    s, i = index(qd)
                    want_i = ident[ref_i[lo:hi]] if use_ident else ref_i[lo:hi]
                    assert np.array_equal(i.cpu().numpy(), want_i), (exchange, use_ident, rep)
                    assert np.array_equal(s.cpu().numpy(), ref_s[lo:hi])
                gs, gi = index.gather(s, i)
                assert np.array_equal(gi.cpu().numpy(), ident[ref_i] if use_ident else ref_i)
        # identical to the single-device search over the concatenated candidates
        s1, i1 = tt.ops.topk_bruteforce("bf16", qd, torch.as_tensor(cand).cuda().to(torch.bfloat16), k)
        assert np.array_equal(i1.cpu().numpy(), ref_i)
        results[rank] = "ok"
    except Exception:
        import traceback
        results[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def test_candidate_sharded_topk_equals_single_device_ids():
    WORLD = _world()
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker_serving, args=(WORLD, _free_port(), results), nprocs=WORLD, join=True)
    for r in range(WORLD):
        assert results.get(r) == "ok", results.get(r)


def _worker_direct(rank, port, results, WORLD=2):
    """The dC pass that scatters its row blocks straight into the owners' slots (TMA stores to peer memory) against the
    combine + scatter kernel: same arithmetic, so losses and shards must be IDENTICAL.  Needs an unsplit dC pass:
    world * b >= 128 * #SMs."""
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=torch.device("cuda", rank))
    try:
        import two_tower_b200 as tt
        from two_tower_b200 import parallel, synth
        tt.set_precision("bf16")
        sms = torch.cuda.get_device_properties(rank).multi_processor_count
        b = -(-64 * sms // 128) * 128                                       # 2 * b >= 128 * SMs, multiple of 128
        cfg = synth.Config("big", 78, b, 128, 40001, 30011, (256, 128), 0.2)
        out = {}
        for mode in ("1", "0"):
            os.environ["TT_DC_DIRECT"] = mode
            tt.core.config.seed = 1234
            torch.manual_seed(7)
            model = parallel.build_sharded_two_tower(cfg, dist.group.WORLD, lr=0.05, peer="exchange")
            model.test_step(synth.make_batch(cfg, 10 + rank))
            for ti, seq in enumerate((model.user_model, model.item_model)):
                for li, l in enumerate(seq.layers[1:]):
                    for vi, v in enumerate((l.kernel, l.bias)):
                        g = torch.Generator(device="cuda"); g.manual_seed(100 * ti + 10 * li + vi)
                        v.value.copy_(torch.randn(v.value.shape, device="cuda", generator=g) * 0.05)
                        v.refresh_shadows()
            losses = [float(model.train_step(synth.make_batch(cfg, 100 + 10 * s + rank))["loss"].item()) for s in range(3)]
            assert (model.exchange.dc_maps is not None) == (mode == "1")
            out[mode] = (losses, model.item_model.layers[0].embeddings.value.clone(), model.item_model.layers[1].kernel.value.clone())
        assert out["1"][0] == out["0"][0], (out["1"][0], out["0"][0])
        assert torch.equal(out["1"][1], out["0"][1]) and torch.equal(out["1"][2], out["0"][2])
        assert all(np.isfinite(x) for x in out["1"][0])
        results[rank] = "ok"
    except Exception:
        import traceback
        results[rank] = traceback.format_exc()
    finally:
        os.environ.pop("TT_DC_DIRECT", None)
        dist.destroy_process_group()


def test_dc_pass_scattering_into_owner_slots_is_identical_to_combine_scatter():
    if os.environ.get("TT_TEST_DC_DIRECT") != "1":
        pytest.skip("experimental path (TT_DC_DIRECT=1), not enabled by default: set TT_TEST_DC_DIRECT=1 to run")
    WORLD = 2
    if torch.cuda.device_count() < WORLD:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker_direct, args=(_free_port(), results), nprocs=WORLD, join=True)
    for r in range(WORLD):
        assert results.get(r) == "ok", results.get(r)
