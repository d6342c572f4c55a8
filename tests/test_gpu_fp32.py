"""GPU parity, fp32 precision: every kernel through the C-ABI (two_tower_b200.ops) against the
oracle on the same seeded inputs.  Integer/index results bit-exact; floats within 1e-5
relative (north_star tolerance for fp32)."""
import numpy as np
import pytest
import torch

import oracle
from two_tower_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def dev(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="module")
def ops(tt):
    tt.ops.device_check()
    return tt.ops


# ---------------------------------------------------------------------------------- K1
class TestTowerInput:
    @pytest.mark.parametrize("n,d,V", [(1, 4, 3), (257, 64, 1000), (8192, 128, 50000), (33, 200, 77), (5, 1024, 9)])
    def test_gather_rows_bit_exact(self, ops, n, d, V):
        rng = synth.rng_for(n + d)
        table = oracle.keras_uniform(rng, (V, d))
        ids = synth.draw_ids(rng, n, V, zipf=1.2 if n > 100 else None)
        out = ops.embedding_gather(dev(table), dev(ids)).cpu().numpy()
        assert np.array_equal(out, oracle.embedding_lookup(table, ids))
        out16 = ops.embedding_gather(dev(table), dev(ids), torch.bfloat16).float().cpu().numpy()
        assert np.array_equal(out16, oracle.bf16_round(table[ids]))

    def test_empty_batch(self, ops):
        table = dev(np.zeros((4, 8), np.float32))
        out = ops.embedding_gather(table, dev(np.zeros(0, np.int64)))
        assert out.shape == (0, 8)

    @pytest.mark.parametrize("mode", ["sum", "mean"])
    def test_bag_pooling_bit_exact_with_empty_and_long_bags(self, ops, mode):
        rng = synth.rng_for(17)
        table = oracle.keras_uniform(rng, (500, 64))
        vals, offs = synth.draw_bags(rng, 300, 500, 0, 9)
        lens = np.diff(offs)
        assert (lens == 0).any()
        # one long bag (> 32 members) to cross the 32-id strip boundary
        long_vals = synth.draw_ids(rng, 70, 500)
        vals = np.concatenate([vals, long_vals]); offs = np.concatenate([offs, [offs[-1] + 70]])
        out = ops.embedding_bag(dev(table), dev(vals), dev(offs), mode).cpu().numpy()
        ref = oracle.embedding_bag(table, vals, offs, mode, dtype=np.float32)
        assert np.array_equal(out, ref)
        assert np.array_equal(out[np.flatnonzero(np.diff(offs) == 0)], np.zeros(((np.diff(offs) == 0).sum(), 64), np.float32))

    def test_golden_bag(self, ops, golden_dir):
        g = np.load(golden_dir / "cfg1.npz")
        cfg = synth.CONFIGS["cfg1"]; rng = synth.rng_for(cfg.seed)
        oracle.keras_uniform(rng, (cfg.v_user, cfg.dim)); I = oracle.keras_uniform(rng, (cfg.v_item, cfg.dim))
        out = ops.embedding_bag(dev(I), dev(g["bag_values"]), dev(g["bag_offsets"]), "mean").cpu().numpy()
        assert rel_err(out, g["bag_mean"]) < 1e-6

    def test_fused_three_feature_tower_input(self, ops):
        rng = synth.rng_for(5)
        t_id, t_cat, t_br = (oracle.keras_uniform(rng, (v, 128)) for v in (1000, 64, 300))
        ids = synth.draw_ids(rng, 777, 1000)
        cat = synth.draw_bags(rng, 777, 64, 1, 8)
        br = synth.draw_bags(rng, 777, 300, 0, 2)
        f32, b16 = ops.tower_input_fwd([(dev(t_id), dev(ids), None, "sum"), (dev(t_cat), dev(cat[0]), dev(cat[1]), "mean"),
                                        (dev(t_br), dev(br[0]), dev(br[1]), "sum")], 777, 128, True, True)
        ref = t_id[ids]
        ref = ref + oracle.embedding_bag(t_cat, cat[0], cat[1], "mean", np.float32)
        ref = ref + oracle.embedding_bag(t_br, br[0], br[1], "sum", np.float32)
        assert np.array_equal(f32.cpu().numpy(), ref)
        assert np.array_equal(b16.float().cpu().numpy(), oracle.bf16_round(ref))

    def test_out_of_range_id_raises_fault_flag(self, ops):
        table = dev(np.ones((10, 8), np.float32))
        flag = torch.zeros(1, dtype=torch.int32, device="cuda")
        ops.tower_input_fwd([(table, dev(np.array([1, 10, 2], np.int64)), None, "sum")], 3, 8, True, False, flag)
        assert int(flag.item()) == 1


# ---------------------------------------------------------------------------------- K5
class TestSparseOptimizer:
    def _run(self, ops, V, d, ids, rows, offs=None, mode="sum", lr=0.1, steps=2):
        rng = synth.rng_for(V + d)
        table = oracle.keras_uniform(rng, (V, d)); acc = np.full((V, d), 0.1, np.float32)
        t_d, a_d = dev(table), dev(acc)
        nnz = len(ids)
        ws = ops.SparseWorkspace(nnz, d, t_d.device)
        flag = torch.zeros(nnz, dtype=torch.uint8, device="cuda")
        t_ref, a_ref = table.astype(np.float64), acc.astype(np.float64)
        for _ in range(steps):       # second step proves the workspace was left clean
            ops.sparse_adagrad_update(t_d, a_d, dev(ids), None if offs is None else dev(offs), mode, dev(rows), lr, 1e-7, ws, flag)
            if offs is None:
                e_ids, e_rows = ids, rows
            else:
                e_ids, e_rows = oracle.embedding_bag_backward(ids, offs, rows, mode)
            t_ref, a_ref, uniq = oracle.adagrad_sparse(t_ref, a_ref, e_ids, e_rows.astype(np.float64), lr, 1e-7)
            got_unique = ids[flag.cpu().numpy().astype(bool)]
            assert np.array_equal(got_unique, uniq)                      # tf.unique order, bit-exact row set
            t_out, a_out = t_d.cpu().numpy(), a_d.cpu().numpy()
            assert rel_err(t_out[uniq], t_ref[uniq]) < RTOL and rel_err(a_out[uniq], a_ref[uniq]) < RTOL
            untouched = np.setdiff1d(np.arange(V), uniq)
            assert np.array_equal(t_out[untouched], table[untouched])    # untouched rows bit-identical
            assert np.array_equal(a_out[untouched], acc[untouched])

    def test_ids_with_heavy_duplicates(self, ops):
        rng = synth.rng_for(1)
        ids = synth.draw_ids(rng, 4096, 1000, zipf=1.1)
        self._run(ops, 1000, 64, ids, rng.normal(size=(4096, 64)).astype(np.float32) * 0.01)

    def test_all_unique_and_odd_dim(self, ops):
        rng = synth.rng_for(2)
        ids = rng.permutation(5000)[:777].astype(np.int64)
        self._run(ops, 5000, 36, ids, rng.normal(size=(777, 36)).astype(np.float32))

    @pytest.mark.parametrize("mode", ["sum", "mean"])
    def test_bag_gradients(self, ops, mode):
        rng = synth.rng_for(3)
        vals, offs = synth.draw_bags(rng, 500, 200, 0, 6)
        self._run(ops, 200, 32, vals, rng.normal(size=(500, 32)).astype(np.float32), offs, mode)

    def test_one_workspace_serves_batches_of_different_sizes(self, ops):
        """Ragged bags change nnz from step to step while the caller keeps ONE workspace (sized for the largest batch).
        The workspace layout must not move with nnz: its accumulation rows are only zero where tt_sparse_workspace_init
        (and the self-cleaning launches) left them.  The allocator hands out DIRTY memory here on purpose."""
        rng = synth.rng_for(77)
        V, d, lr = 60, 32, 0.1
        table = oracle.keras_uniform(rng, (V, d)); acc = np.full((V, d), 0.1, np.float32)
        t_d, a_d = dev(table), dev(acc)
        nbytes = ops.SparseWorkspace(700, d, t_d.device).nbytes
        junk = torch.full((4 * nbytes,), 0x7F, dtype=torch.uint8, device="cuda")      # NaN-ish garbage in the freed blocks
        del junk
        ws = ops.SparseWorkspace(700, d, t_d.device)
        t_ref, a_ref = table.astype(np.float64), acc.astype(np.float64)
        for rows_n, lo, hi in ((160, 0, 8), (96, 0, 4), (200, 1, 3), (33, 0, 9), (96, 0, 4)):    # nnz ~ 640, 190, 400, 150, 190
            vals, offs = synth.draw_bags(rng, rows_n, V, lo, hi)
            g = rng.normal(size=(rows_n, d)).astype(np.float32)
            ops.sparse_adagrad_update(t_d, a_d, dev(vals), dev(offs), "mean", dev(g), lr, 1e-7, ws)
            e_ids, e_rows = oracle.embedding_bag_backward(vals, offs, g, "mean")
            t_ref, a_ref, _ = oracle.adagrad_sparse(t_ref, a_ref, e_ids, e_rows.astype(np.float64), lr, 1e-7)
            assert rel_err(t_d.cpu().numpy(), t_ref) < RTOL and rel_err(a_d.cpu().numpy(), a_ref) < RTOL

    def test_golden_adagrad(self, ops, golden_dir):
        g = np.load(golden_dir / "cfg1.npz")
        cfg = synth.CONFIGS["cfg1"]; rng = synth.rng_for(cfg.seed)
        U = oracle.keras_uniform(rng, (cfg.v_user, cfg.dim))
        t_d, a_d = dev(U), dev(np.full(U.shape, 0.1, np.float32))
        ws = ops.SparseWorkspace(cfg.batch, cfg.dim, t_d.device)
        flag = torch.zeros(cfg.batch, dtype=torch.uint8, device="cuda")
        ops.sparse_adagrad_update(t_d, a_d, dev(g["uid"]), None, "sum", dev(g["temp_dq"]), 0.1, 1e-7, ws, flag)
        uniq = g["adagrad_unique_ids"]
        assert np.array_equal(g["uid"][flag.cpu().numpy().astype(bool)], uniq)
        assert rel_err(t_d.cpu().numpy()[uniq], g["adagrad_rows"]) < RTOL
        assert rel_err(a_d.cpu().numpy()[uniq], g["adagrad_acc_rows"]) < RTOL

    def test_lazy_adam(self, ops):
        rng = synth.rng_for(4)
        V, d = 300, 16
        T = oracle.keras_uniform(rng, (V, d)); m = np.zeros((V, d), np.float32); v = np.zeros((V, d), np.float32)
        ids = synth.draw_ids(rng, 128, V, zipf=1.3); rows = rng.normal(size=(128, d)).astype(np.float32)
        t_d, m_d, v_d = dev(T), dev(m), dev(v)
        ws = ops.SparseWorkspace(128, d, t_d.device)
        t_ref, m_ref, v_ref = T.astype(np.float64), m.astype(np.float64), v.astype(np.float64)
        for step in (1, 2):
            alpha = 0.01 * np.sqrt(1 - 0.999 ** step) / (1 - 0.9 ** step)
            ops.sparse_lazy_adam_update(t_d, m_d, v_d, dev(ids), None, "sum", dev(rows), float(alpha), 0.9, 0.999, 1e-7, ws)
            t_ref, m_ref, v_ref, _ = oracle.lazy_adam_sparse(t_ref, m_ref, v_ref, ids, rows.astype(np.float64), step, lr=0.01)
        assert rel_err(t_d.cpu().numpy(), t_ref) < RTOL and rel_err(v_d.cpu().numpy(), v_ref) < RTOL

    def test_dense_adagrad_and_adam_with_parts_and_l2(self, ops):
        rng = synth.rng_for(6)
        w = rng.normal(size=(24, 40)).astype(np.float32); parts = rng.normal(size=(3, 24, 40)).astype(np.float32)
        w_d, a_d = dev(w), dev(np.full_like(w, 0.1))
        sh = torch.empty((24, 40), dtype=torch.bfloat16, device="cuda")
        ops.dense_adagrad_update(w_d, a_d, dev(parts), 3, 0.05, 1e-7, 1e-3, sh)
        g = parts.astype(np.float64).sum(0) + 2e-3 * w
        w_ref, a_ref = oracle.adagrad_dense(w.astype(np.float64), np.full(w.shape, 0.1), g, 0.05, 1e-7)
        assert rel_err(w_d.cpu().numpy(), w_ref) < RTOL and rel_err(a_d.cpu().numpy(), a_ref) < RTOL
        assert np.array_equal(sh.float().cpu().numpy(), oracle.bf16_round(w_d.cpu().numpy()))
        m_d, v_d, w2 = dev(np.zeros_like(w)), dev(np.zeros_like(w)), dev(w)
        alpha = 0.001 * np.sqrt(1 - 0.999) / (1 - 0.9)
        ops.dense_adam_update(w2, m_d, v_d, dev(parts[:1]), 1, float(alpha), 0.9, 0.999, 1e-7)
        w_ref, _, _ = oracle.adam_dense(w.astype(np.float64), 0 * w, 0 * w, parts[0].astype(np.float64), 1)
        assert rel_err(w2.cpu().numpy(), w_ref) < RTOL


# ---------------------------------------------------------------------------------- K2
class TestDenseFp32:
    @pytest.mark.parametrize("M,i,o", [(256, 64, 32), (1000, 128, 256), (77, 20, 36)])
    def test_forward_backward(self, ops, M, i, o):
        rng = synth.rng_for(M)
        x = np.maximum(rng.normal(size=(M, i)), 0).astype(np.float32)     # looks like a relu output
        k = oracle.glorot_uniform(rng, i, o); b = rng.normal(size=o).astype(np.float32) * 0.1
        y, _ = ops.dense_fwd("fp32", dev(x), dev(k), dev(b), relu=True)
        ref = oracle.dense_forward(x.astype(np.float64), k, b, "relu")
        assert rel_err(y.cpu().numpy(), ref) < RTOL
        dy = rng.normal(size=(M, o)).astype(np.float32)
        dx, _, dk, P, db = ops.dense_bwd("fp32", dev(dy), dev(x), dev(k), relu_mask_x=True, want_dx=True)
        assert P == 1
        assert rel_err(dx.cpu().numpy(), (dy.astype(np.float64) @ k.T) * (x > 0)) < RTOL
        assert rel_err(dk[0].cpu().numpy(), x.astype(np.float64).T @ dy) < RTOL
        assert rel_err(db.sum(0).cpu().numpy(), dy.astype(np.float64).sum(0)) < RTOL


# ------------------------------------------------------------------------------- K3/K4
class TestRetrievalFp32:
    def _check(self, ops, q, c, label_offset=0, **kw):
        T = kw.get("temperature")
        inv_t = 1.0 if T is None else 1.0 / T
        w, p, ids = kw.get("sample_weight"), kw.get("candidate_sampling_probability"), kw.get("candidate_ids")
        logq = None if p is None else dev(np.log(np.clip(p, 1e-6, 1.0)), torch.float32)
        ids_d = dev(ids) if kw.get("remove_accidental_hits") else None
        w_d = None if w is None else dev(w, torch.float32)
        nq = q.shape[0]
        # oracle with the positives at label_offset: rotate the candidates so eye labels apply
        perm = np.concatenate([np.arange(label_offset, label_offset + nq), np.arange(0, label_offset),
                               np.arange(label_offset + nq, c.shape[0])])
        kw_o = dict(kw)
        if p is not None: kw_o["candidate_sampling_probability"] = np.asarray(p)[perm]
        if ids is not None: kw_o["candidate_ids"] = np.asarray(ids)[perm]
        r = oracle.retrieval_loss_and_grads(q.astype(np.float64), c.astype(np.float64)[perm], **kw_o)
        loss, lse, pos = ops.retrieval_loss_fwd("fp32", dev(q), dev(c), inv_t, label_offset, w_d, logq, ids_d)
        assert float(loss.item()) == pytest.approx(r["loss"], rel=RTOL)
        assert rel_err(lse.cpu().numpy(), r["lse"]) < RTOL and rel_err(pos.cpu().numpy(), r["pos"]) < RTOL
        g = ops.retrieval_loss_bwd("fp32", dev(q), dev(c), inv_t, lse, label_offset, w_d, logq, ids_d)
        dc_ref = np.empty_like(r["dc"]); dc_ref[perm] = r["dc"]
        assert rel_err(g["dq"].cpu().numpy(), r["dq"]) < RTOL
        assert rel_err(g["dc"].cpu().numpy(), dc_ref) < RTOL
        return r

    @pytest.mark.parametrize("tag", ["plain", "temp", "full"])
    def test_golden_cfg1(self, ops, golden_dir, tag):
        g = np.load(golden_dir / "cfg1.npz")
        cfg = synth.CONFIGS["cfg1"]; rng = synth.rng_for(cfg.seed)
        U = oracle.keras_uniform(rng, (cfg.v_user, cfg.dim)); I = oracle.keras_uniform(rng, (cfg.v_item, cfg.dim))
        kw = {"plain": {}, "temp": dict(temperature=cfg.temperature),
              "full": dict(temperature=cfg.temperature, sample_weight=g["w"], candidate_sampling_probability=g["p"],
                           candidate_ids=g["iid"], remove_accidental_hits=True)}[tag]
        r = self._check(ops, U[g["uid"]], I[g["iid"]], **kw)
        assert r["loss"] == pytest.approx(float(g[f"{tag}_loss"]), rel=1e-9)

    @pytest.mark.parametrize("nq,nc,d,off", [(1, 1, 4, 0), (100, 173, 20, 0), (64, 64, 128, 0), (130, 600, 256, 257), (300, 1200, 64, 900)])
    def test_ragged_shapes_and_label_offset(self, ops, nq, nc, d, off):
        rng = synth.rng_for(nq * 7 + nc)
        q = rng.normal(size=(nq, d)).astype(np.float32) * 0.3; c = rng.normal(size=(nc, d)).astype(np.float32) * 0.3
        self._check(ops, q, c, off, temperature=0.5, sample_weight=rng.uniform(0.5, 1.5, nq))

    def test_full_size_property_uniform_embeddings(self, ops):
        # B=8192, d=128: identical rows -> every logit equal -> loss = B*log(B), dq = 0
        B, d = 8192, 128
        q = torch.full((B, d), 0.05, device="cuda"); c = torch.full((B, d), 0.05, device="cuda")
        loss, lse, _ = ops.retrieval_loss_fwd("fp32", q, c, 10.0)
        assert float(loss.item()) == pytest.approx(B * np.log(B), rel=1e-5)
        g = ops.retrieval_loss_bwd("fp32", q, c, 10.0, lse)
        assert float(g["dq"].abs().max().item()) < 1e-6
        assert float(g["dc"].abs().max().item()) < 1e-5    # each column of (softmax - eye) sums to 0


# ---------------------------------------------------------------------------------- K6
def check_topk(scores, ids, q, c, k, identifiers=None, exact=True):
    """exact: dyadic data, every sum exact -> equal to the fp64 oracle.  Otherwise (real-valued data) the order is that of
    tf.math.top_k on the correctly rounded fp32 score tensor (exact re-rank, csrc/topk_rerank.cu): still bit-exact."""
    ref_s, ref_i = oracle.brute_force_topk(q, c, k, identifiers, score_dtype=None if exact else np.float32)
    ids, scores = ids.cpu().numpy(), scores.cpu().numpy()
    assert np.array_equal(ids, ref_i)
    assert np.array_equal(scores.astype(ref_s.dtype), ref_s)


class TestTopKFp32:
    @pytest.mark.parametrize("nq,nc,d,k", [(70, 1000, 64, 100), (5, 37, 8, 37), (200, 5000, 128, 10), (64, 300, 256, 64), (3, 100, 16, 1)])
    def test_exact_arithmetic_data_bit_exact_ids_with_ties(self, ops, nq, nc, d, k):
        rng = synth.rng_for(nq + nc + k)
        q, c = synth.exact_matrix(rng, nq, d, 2), synth.exact_matrix(rng, nc, d, 2)
        s, i = ops.topk_bruteforce("fp32", dev(q), dev(c), k)
        check_topk(s, i, q, c, k)

    def test_split_candidates_path_and_identifiers(self, ops, tt):
        rng = synth.rng_for(99)
        nq, nc, d, k = 40, 50000, 32, 100
        assert tt._lib.load().tt_topk_num_splits(0, nq, nc, d, k) > 1
        q, c = synth.exact_matrix(rng, nq, d, 3), synth.exact_matrix(rng, nc, d, 3)
        ident = rng.permutation(nc).astype(np.int64) + 10_000_000_000
        s, i = ops.topk_bruteforce("fp32", dev(q), dev(c), k, identifiers=dev(ident))
        check_topk(s, i, q, c, k, ident)
        s, i = ops.topk_bruteforce("fp32", dev(q), dev(c), k, cand_index_base=12345)
        check_topk(s, i - 12345, q, c, k)

    def test_gaussian_and_golden(self, ops, golden_dir):
        rng = synth.rng_for(7)
        q = (rng.normal(size=(300, 128)) / np.sqrt(128)).astype(np.float32)
        c = (rng.normal(size=(20000, 128)) / np.sqrt(128)).astype(np.float32)
        s, i = ops.topk_bruteforce("fp32", dev(q), dev(c), 100)
        check_topk(s, i, q, c, 100, exact=False)
        g = np.load(golden_dir / "cfg1.npz")
        cfg = synth.CONFIGS["cfg1"]; r2 = synth.rng_for(cfg.seed)
        U = oracle.keras_uniform(r2, (cfg.v_user, cfg.dim)); I = oracle.keras_uniform(r2, (cfg.v_item, cfg.dim))
        s, i = ops.topk_bruteforce("fp32", dev(U[g["uid"]]), dev(I), 100)
        # golden ids were frozen from the fp64-ordered oracle: equal unless two fp64 scores round to one fp32
        ref_s, ref_i = oracle.brute_force_topk(U[g["uid"]], I, 100, score_dtype=np.float32)
        assert np.array_equal(i.cpu().numpy(), ref_i)
        assert (i.cpu().numpy() != g["topk_ids"]).mean() < 1e-3
        assert rel_err(s.cpu().numpy(), g["topk_scores"]) < 1e-6

    def test_merge_lists(self, ops):
        rng = synth.rng_for(8)
        q, c = synth.exact_matrix(rng, 33, 16, 2), synth.exact_matrix(rng, 800, 16, 2)
        parts = [oracle.brute_force_topk(q, c[lo:lo + 100], 20, identifiers=np.arange(lo, lo + 100)) for lo in range(0, 800, 100)]
        S = np.stack([p[0] for p in parts]).astype(np.float32); I = np.stack([p[1] for p in parts]).astype(np.int64)
        s, i = ops.topk_merge(dev(S), dev(I), 20)
        ref_s, ref_i = oracle.brute_force_topk(q, c, 20)
        assert np.array_equal(i.cpu().numpy(), ref_i) and np.array_equal(s.cpu().numpy().astype(np.float64), ref_s)

    def test_factorized_topk_metric_both_modes(self, tt):
        tt.set_precision("fp32")
        rng = synth.rng_for(10)
        cands = (rng.normal(size=(3000, 32))).astype(np.float32)
        q = rng.normal(size=(500, 32)).astype(np.float32)
        true_idx = rng.integers(0, 3000, 500)
        ks = (1, 5, 10, 20, 50, 100)
        w = rng.uniform(0.5, 2.0, 500)
        for ids_mode in (False, True):
            m = tt.metrics.FactorizedTopK(torch.as_tensor(cands).cuda(), ks=ks)
            o = oracle.FactorizedTopKOracle(cands, ks=ks)
            for lo in (0, 250):
                sl = slice(lo, lo + 250)
                kw = dict(true_candidate_ids=true_idx[sl]) if ids_mode else {}
                m.update_state(torch.as_tensor(q[sl]).cuda(), torch.as_tensor(cands[true_idx[sl]]).cuda(), sample_weight=w[sl], **kw)
                o.update_state(q[sl], cands[true_idx[sl]], sample_weight=w[sl], **kw)
            got, ref = m.result(), o.result()
            for k in ks:
                name = f"factorized_top_k/top_{k}_categorical_accuracy"
                assert got[name] == pytest.approx(ref[name], abs=2e-3)


# ---------------------------------------------------------------- whole train step (a7)
class TestTrainStepFp32:
    def _build(self, tt, vu, vi, vc, d, mlp, temperature, l2=0.0, lr=0.1):
        tt.set_precision("fp32")

        class TwoTower(tt.models.Model):
            def __init__(s):
                super().__init__()
                s.user_model = tt.Sequential([tt.layers.Embedding(vu, d)] + [tt.layers.Dense(u, "relu" if j < len(mlp) - 1 else None, kernel_regularizer=l2) for j, u in enumerate(mlp)])
                s.item_in = tt.FeatureSum({"item_id_encoded": tt.layers.Embedding(vi, d), "category": tt.layers.EmbeddingBag(vc, d, "mean")})
                s.item_mlp = tt.Sequential([tt.layers.Dense(u, "relu" if j < len(mlp) - 1 else None, kernel_regularizer=l2) for j, u in enumerate(mlp)])
                s.task = tt.tasks.Retrieval(temperature=temperature)

            def compute_loss(s, features, training=False):
                return s.task(s.user_model(features["user_id_encoded"]), s.item_mlp(s.item_in(features)))

        m = TwoTower()
        m.compile(optimizer=tt.optimizers.Adagrad(lr))
        return m

    def test_three_steps_match_oracle(self, tt):
        vu, vi, vc, d, mlp, T, l2, lr = 300, 200, 20, 32, (48, 16), 0.2, 1e-4, 0.1
        model = self._build(tt, vu, vi, vc, d, mlp, T, l2, lr)
        rng = synth.rng_for(42)
        qs = oracle.TowerSpec([("user_id_encoded", "id", vu, None)], d, mlp)
        cs = oracle.TowerSpec([("item_id_encoded", "id", vi, None), ("category", "bag", vc, "mean")], d, mlp)
        qp, cp = oracle.init_tower(qs, rng, np.float32), oracle.init_tower(cs, rng, np.float32)
        # first call builds the Dense layers; then load the oracle's initial weights
        batch0 = {"user_id_encoded": synth.draw_ids(rng, 96, vu, 1.2), "item_id_encoded": synth.draw_ids(rng, 96, vi, 1.2),
                  "category": synth.draw_bags(rng, 96, vc, 0, 4)}
        model.test_step(batch0)
        model.user_model.layers[0].set_weights([qp["tables"]["user_id_encoded"]])
        model.item_in.features["item_id_encoded"].set_weights([cp["tables"]["item_id_encoded"]])
        model.item_in.features["category"].set_weights([cp["tables"]["category"]])
        for j in range(len(mlp)):
            model.user_model.layers[1 + j].set_weights([qp["kernels"][j], qp["biases"][j]])
            model.item_mlp.layers[j].set_weights([cp["kernels"][j], cp["biases"][j]])
        mk = lambda p: {"tables": {k: np.full(v.shape, 0.1) for k, v in p["tables"].items()},
                        "kernels": [np.full(k.shape, 0.1) for k in p["kernels"]], "biases": [np.full(b.shape, 0.1) for b in p["biases"]]}
        qsl, csl = mk(qp), mk(cp)
        qp = {"tables": {k: v.astype(np.float64) for k, v in qp["tables"].items()}, "kernels": [k.astype(np.float64) for k in qp["kernels"]], "biases": [b.astype(np.float64) for b in qp["biases"]]}
        cp = {"tables": {k: v.astype(np.float64) for k, v in cp["tables"].items()}, "kernels": [k.astype(np.float64) for k in cp["kernels"]], "biases": [b.astype(np.float64) for b in cp["biases"]]}
        for step in range(3):
            b = {"user_id_encoded": synth.draw_ids(rng, 96, vu, 1.2), "item_id_encoded": synth.draw_ids(rng, 96, vi, 1.2),
                 "category": synth.draw_bags(rng, 96, vc, 0, 4)}
            out = model.train_step(b)
            ref = oracle.two_tower_train_step(qs, cs, qp, cp, qsl, csl, {"user_id_encoded": b["user_id_encoded"]},
                                              {"item_id_encoded": b["item_id_encoded"], "category": b["category"]},
                                              temperature=T, lr=lr, l2=l2)
            assert float(out["loss"].item()) == pytest.approx(ref["loss"], rel=RTOL)
            assert float(out["regularization_loss"].item()) == pytest.approx(ref["regularization_loss"], rel=RTOL)
            assert float(out["total_loss"].item()) == pytest.approx(ref["total_loss"], rel=RTOL)
        assert rel_err(model.user_model.layers[0].get_weights()[0], qp["tables"]["user_id_encoded"]) < RTOL
        assert rel_err(model.item_in.features["category"].get_weights()[0], cp["tables"]["category"]) < RTOL
        for j in range(len(mlp)):
            k, bias = model.user_model.layers[1 + j].get_weights()
            assert rel_err(k, qp["kernels"][j]) < 5 * RTOL and rel_err(bias, qp["biases"][j]) < 5 * RTOL
            k, bias = model.item_mlp.layers[j].get_weights()
            assert rel_err(k, cp["kernels"][j]) < 5 * RTOL

    def test_cuda_graph_step_equals_eager(self, tt):
        # well-conditioned regime (T=0.5, lr=0.01): the only run-to-run difference is the fp32
        # atomic order over duplicate ids, which must stay at the 1e-6 level
        model_a = self._build(tt, 500, 400, 30, 64, (64, 32), 0.5, lr=0.01)
        model_b = self._build(tt, 500, 400, 30, 64, (64, 32), 0.5, lr=0.01)
        rng = synth.rng_for(77)
        mkb = lambda: {"user_id_encoded": torch.as_tensor(synth.draw_ids(rng, 128, 500)).cuda(),
                       "item_id_encoded": torch.as_tensor(synth.draw_ids(rng, 128, 400)).cuda(),
                       "category": tuple(torch.as_tensor(a).cuda() for a in synth.draw_bags(rng, 128, 30, 2, 2))}
        b0 = mkb()
        model_a.test_step(b0); model_b.test_step(b0)
        for va, vb in zip(model_a.trainable_variables, model_b.trainable_variables):
            vb.assign(va.numpy())
        graphed = model_b.make_graphed_train_step(b0, warmup=2)
        for _ in range(2):
            model_a.train_step(b0)          # mirror the two warm-up steps (capture itself executes nothing)
        for _ in range(3):
            b = mkb()
            la = float(model_a.train_step(b)["loss"].item())
            lb = float(graphed(b)["loss"].item())
            assert lb == pytest.approx(la, rel=1e-5)

    def test_three_steps_per_execution_equal_three_eager_steps(self, tt):
        """make_graphed_train_step(steps_per_execution=3): one replay = three consecutive train steps on three batches
        (host batches: one H2D copy for all of them); losses and the iteration count follow the eager model."""
        model_a = self._build(tt, 500, 400, 30, 64, (64, 32), 0.5, lr=0.01)
        model_b = self._build(tt, 500, 400, 30, 64, (64, 32), 0.5, lr=0.01)
        rng = synth.rng_for(78)
        mkb = lambda: {"user_id_encoded": torch.as_tensor(synth.draw_ids(rng, 128, 500)),
                       "item_id_encoded": torch.as_tensor(synth.draw_ids(rng, 128, 400)),
                       "category": tuple(torch.as_tensor(a) for a in synth.draw_bags(rng, 128, 30, 2, 2))}
        cuda = lambda b: {k: (tuple(a.cuda() for a in v) if isinstance(v, tuple) else v.cuda()) for k, v in b.items()}
        b0 = mkb()
        model_a.test_step(cuda(b0)); model_b.test_step(cuda(b0))
        for va, vb in zip(model_a.trainable_variables, model_b.trainable_variables):
            vb.assign(va.numpy())
        graphed = model_b.make_graphed_train_step(cuda(b0), warmup=1, steps_per_execution=3)
        model_a.train_step(cuda(b0))        # mirror the warm-up step (capture itself executes nothing)
        it0 = model_b.optimizer.iterations
        for rnd in range(2):
            group = [mkb() for _ in range(3)]
            la = [float(model_a.train_step(cuda(b))["loss"].item()) for b in group]
            outs = graphed(group if rnd == 0 else graphed.pack([cuda(b) for b in group]))   # host batches / packed device batches
            assert len(outs) == 3
            torch.cuda.synchronize()
            for x, o in zip(la, outs):
                assert float(o["loss"].item()) == pytest.approx(x, rel=1e-5)
            # the losses of one execution are also ONE contiguous device tensor (a single D2H copy reads them)
            assert torch.equal(graphed.losses, torch.cat([o["loss"] for o in outs]))
        assert model_b.optimizer.iterations == it0 + 6
        with pytest.raises(ValueError):
            graphed([mkb()])


# ---------------------------------------------------------------- sharding helpers (8e)
class TestShardingHelpers:
    @pytest.mark.parametrize("n,world", [(0, 2), (1, 1), (5000, 2), (8192, 8), (1025, 3)])
    def test_partition_is_stable_and_exact(self, ops, n, world):
        rng = synth.rng_for(n + world)
        ids = synth.draw_ids(rng, n, 10 ** 8)
        send, perm, counts = ops.partition_ids(dev(ids), world)
        owner = ids % world
        order = np.argsort(owner, kind="stable")
        assert np.array_equal(counts.cpu().numpy(), np.bincount(owner, minlength=world))
        assert np.array_equal(send.cpu().numpy(), (ids // world)[order])
        ref_perm = np.empty(n, np.int64); ref_perm[order] = np.arange(n)
        assert np.array_equal(perm.cpu().numpy(), ref_perm)
        if n:
            x = rng.normal(size=(n, 8)).astype(np.float32)
            y = ops.permute_rows(dev(x), perm, inverse=False)
            assert np.array_equal(y.cpu().numpy(), x[order])
            assert np.array_equal(ops.permute_rows(y, perm, inverse=True).cpu().numpy(), x)
            # padded buckets (static all-to-all shapes): bucket o starts at o * cap, padding = -1
            cap = int(np.bincount(owner, minlength=world).max()) + 3
            flag = torch.zeros(1, dtype=torch.int32, device="cuda")
            send_p, perm_p, _ = ops.partition_ids(dev(ids), world, cap, flag)
            ref = np.full(world * cap, -1, np.int64)
            starts = np.concatenate([[0], np.cumsum(np.bincount(owner, minlength=world))[:-1]])
            pos = np.arange(n) - starts[owner[order]] + owner[order] * cap
            ref[pos] = (ids // world)[order]
            assert np.array_equal(send_p.cpu().numpy(), ref) and int(flag.item()) == 0
            xb = torch.as_tensor(x).cuda().to(torch.bfloat16)
            yb = ops.permute_rows(xb, perm_p, inverse=False, out_rows=world * cap, zero_fill=True)
            assert np.array_equal(ops.permute_rows(yb, perm_p, inverse=True).float().cpu().numpy(), xb.float().cpu().numpy())
            if world > 1 and n > 100:
                ops.partition_ids(dev(ids), world, 8, flag)
                assert int(flag.item()) == 1                      # a bucket larger than the capacity is reported
