"""Oracle vs its golden vectors and vs independent invariants (CPU).  The reference pins
nothing on this path (PARITY UNPINNED), so the oracle is checked against (a) the frozen
vectors of tests/golden/cfg1.npz, (b) independent torch CPU routines, (c) finite differences
and algebraic properties."""
import numpy as np
import pytest
import torch

import oracle
from two_tower_b200 import synth


@pytest.fixture(scope="module")
def cfg1(golden_dir):
    g = np.load(golden_dir / "cfg1.npz")
    cfg = synth.CONFIGS["cfg1"]
    rng = synth.rng_for(cfg.seed)
    U = oracle.keras_uniform(rng, (cfg.v_user, cfg.dim))
    I = oracle.keras_uniform(rng, (cfg.v_item, cfg.dim))
    return cfg, g, U, I


class TestGolden:
    def test_inputs_regenerate(self, cfg1):
        cfg, g, U, I = cfg1
        b = synth.make_batch(cfg, 0)
        assert np.array_equal(b["user_id_encoded"], g["uid"])
        assert np.array_equal(b["item_id_encoded"], g["iid"])
        assert len(np.unique(g["uid"])) < cfg.batch          # Zipf ids: duplicates guaranteed

    @pytest.mark.parametrize("tag", ["plain", "temp", "full"])
    def test_retrieval(self, cfg1, tag):
        cfg, g, U, I = cfg1
        q, c = U[g["uid"]].astype(np.float64), I[g["iid"]].astype(np.float64)
        kw = {"plain": {}, "temp": dict(temperature=cfg.temperature),
              "full": dict(temperature=cfg.temperature, sample_weight=g["w"], candidate_sampling_probability=g["p"],
                           candidate_ids=g["iid"], remove_accidental_hits=True)}[tag]
        r = oracle.retrieval_loss_and_grads(q, c, **kw)
        assert r["loss"] == pytest.approx(float(g[f"{tag}_loss"]), rel=1e-12)
        np.testing.assert_allclose(r["lse"], g[f"{tag}_lse"], rtol=1e-12)
        np.testing.assert_allclose(r["dq"], g[f"{tag}_dq"], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(r["dc"], g[f"{tag}_dc"], rtol=1e-6, atol=1e-9)

    def test_fp32_mode_within_1e5(self, cfg1):
        cfg, g, U, I = cfg1
        r = oracle.retrieval_loss_and_grads(U[g["uid"]], I[g["iid"]], temperature=cfg.temperature, dtype=np.float32)
        assert r["loss"] == pytest.approx(float(g["temp_loss"]), rel=1e-5)

    def test_adagrad_and_topk_and_bag(self, cfg1):
        cfg, g, U, I = cfg1
        q, c = U[g["uid"]].astype(np.float64), I[g["iid"]].astype(np.float64)
        r = oracle.retrieval_loss_and_grads(q, c, temperature=cfg.temperature)
        t1, a1, uniq = oracle.adagrad_sparse(U.astype(np.float64), np.full(U.shape, 0.1), g["uid"], r["dq"], lr=0.1)
        assert np.array_equal(uniq, g["adagrad_unique_ids"])
        np.testing.assert_allclose(t1[uniq], g["adagrad_rows"], rtol=1e-6, atol=1e-9)
        s, ids = oracle.brute_force_topk(q, I.astype(np.float64), 100)
        assert np.array_equal(ids, g["topk_ids"])
        pooled = oracle.embedding_bag(I, g["bag_values"], g["bag_offsets"], "mean")
        np.testing.assert_allclose(pooled, g["bag_mean"], rtol=1e-6, atol=1e-9)


class TestRetrievalSemantics:
    def setup_method(self):
        rng = np.random.default_rng(7)
        self.q = rng.normal(size=(17, 8))
        self.c = rng.normal(size=(23, 8))          # extra negatives appended: nc > nq

    def test_matches_torch_cross_entropy(self):
        s = torch.tensor(self.q @ self.c.T / 0.5)
        ref = torch.nn.functional.cross_entropy(s, torch.arange(17), reduction="sum").item()
        assert oracle.retrieval_loss(self.q, self.c, temperature=0.5) == pytest.approx(ref, rel=1e-12)

    def test_loss_nonnegative_and_softmax_normalised(self):
        r = oracle.retrieval_loss_and_grads(self.q, self.c)
        assert r["loss"] >= 0
        p = np.exp(r["scores"] - r["lse"][:, None])
        np.testing.assert_allclose(p.sum(1), 1.0, rtol=1e-12)

    def test_gradients_finite_difference(self):
        kw = dict(temperature=0.3, sample_weight=np.linspace(0.5, 2, 17))
        r = oracle.retrieval_loss_and_grads(self.q, self.c, **kw)
        eps = 1e-6
        for (arr, grad, idx) in ((self.q, r["dq"], (3, 2)), (self.c, r["dc"], (20, 5)), (self.c, r["dc"], (4, 1))):
            a = arr.copy(); a[idx] += eps
            b = arr.copy(); b[idx] -= eps
            fa = oracle.retrieval_loss(a if arr is self.q else self.q, a if arr is self.c else self.c, **kw)
            fb = oracle.retrieval_loss(b if arr is self.q else self.q, b if arr is self.c else self.c, **kw)
            assert (fa - fb) / (2 * eps) == pytest.approx(grad[idx], rel=1e-5, abs=1e-8)

    def test_sampling_probability_is_clipped_logq(self):
        p = np.full(23, 1e-9); p[0] = 2.0
        s0 = oracle.retrieval_scores(self.q, self.c)
        s1 = oracle.retrieval_scores(self.q, self.c, candidate_sampling_probability=p)
        np.testing.assert_allclose(s0[:, 1] - s1[:, 1], np.log(1e-6))
        np.testing.assert_allclose(s0[:, 0] - s1[:, 0], 0.0, atol=1e-15)

    def test_accidental_hits_masked_not_positive(self):
        ids = np.arange(23); ids[5] = ids[2]            # candidate 5 duplicates query 2's positive
        s = oracle.retrieval_scores(self.q, self.c, candidate_ids=ids, remove_accidental_hits=True)
        s0 = oracle.retrieval_scores(self.q, self.c)
        assert s[2, 5] < -1e30 and s[5, 2] < -1e30
        assert s[2, 2] == s0[2, 2] and s[5, 5] == s0[5, 5]
        with pytest.raises(ValueError):
            oracle.retrieval_scores(self.q, self.c, remove_accidental_hits=True)

    def test_every_option_together_matches_torch_autograd(self):
        """tfrs.tasks.Retrieval.call with temperature, sample weights, candidate_sampling_probability (clipped logQ) and
        accidental-hit removal at once, restated in torch (fp64) with autograd for dq / dc; a rectangular score matrix
        (more candidates than queries, as with global-batch negatives) and candidate ids with duplicates of positives."""
        rng = np.random.default_rng(17)
        nq, nc, d, T = 9, 14, 6, 0.2
        q, c = rng.normal(size=(nq, d)), rng.normal(size=(nc, d))
        ids = rng.integers(0, 6, size=nc)                       # many duplicates, also among the positives ids[:nq]
        p = rng.uniform(1e-8, 0.4, size=nc); p[3] = 5e-7        # below the 1e-6 clip
        w = rng.uniform(0.2, 2.0, size=nq)
        r = oracle.retrieval_loss_and_grads(q, c, temperature=T, sample_weight=w, candidate_sampling_probability=p,
                                            candidate_ids=ids, remove_accidental_hits=True)
        tq, tc = torch.tensor(q, requires_grad=True), torch.tensor(c, requires_grad=True)
        s = tq @ tc.T / T - torch.log(torch.clamp(torch.tensor(p), 1e-6, 1.0))[None, :]
        labels = torch.eye(nq, nc, dtype=torch.float64)
        dup = (torch.tensor(ids)[:nq, None] == torch.tensor(ids)[None, :]).double() - labels
        s = s + dup * (float(np.finfo(np.float32).min) / 100.0)
        per_row = -(labels * torch.log_softmax(s, dim=1)).sum(1)
        loss = (torch.tensor(w) * per_row).sum()
        loss.backward()
        assert (dup.sum(1) > 0).any()                           # accidental hits are present
        assert r["loss"] == pytest.approx(float(loss.detach()), rel=1e-12)
        np.testing.assert_allclose(r["dq"], tq.grad.numpy(), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(r["dc"], tc.grad.numpy(), rtol=1e-9, atol=1e-12)

    def test_hard_negative_mining_matches_torch_autograd(self):
        """tfrs.layers.loss.HardNegativeMining(n): per row the positive and the n highest-scoring negatives survive
        (top_k of scores + labels * MAX_FLOAT, k = n + 1), the loss is the cross entropy over the gathered columns."""
        rng = np.random.default_rng(23)
        nq, nc, d, T, n = 8, 20, 5, 0.3, 4
        q, c = rng.normal(size=(nq, d)), rng.normal(size=(nc, d))
        r = oracle.retrieval_loss_and_grads(q, c, temperature=T, num_hard_negatives=n)
        tq, tc = torch.tensor(q, requires_grad=True), torch.tensor(c, requires_grad=True)
        s = tq @ tc.T / T
        labels = torch.eye(nq, nc, dtype=torch.float64)
        _, idx = torch.topk(s.detach() + labels * 1e30, n + 1, dim=1)
        gs, gl = torch.gather(s, 1, idx), torch.gather(labels, 1, idx)
        assert (gl.sum(1) == 1).all()                          # the positive is among the survivors of every row
        loss = -(gl * torch.log_softmax(gs, dim=1)).sum()
        loss.backward()
        assert r["loss"] == pytest.approx(float(loss.detach()), rel=1e-12)
        np.testing.assert_allclose(r["dq"], tq.grad.numpy(), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(r["dc"], tc.grad.numpy(), rtol=1e-9, atol=1e-12)

    def test_hard_negatives_keep_positive_plus_n(self):
        r = oracle.retrieval_loss_and_grads(self.q, self.c, num_hard_negatives=3)
        assert ((r["dq"] != 0).any())
        full = oracle.retrieval_loss(self.q, self.c)
        assert r["loss"] <= full + 1e-12                 # fewer negatives -> smaller logsumexp


class TestTowersAndOptimizers:
    def test_bag_matches_torch_embedding_bag_and_backward(self):
        rng = np.random.default_rng(3)
        T = rng.normal(size=(50, 6))
        vals, offs = synth.draw_bags(rng, 12, 50, 0, 5)
        for mode in ("sum", "mean"):
            out = oracle.embedding_bag(T, vals, offs, mode)
            tt_ = torch.tensor(T, requires_grad=True)
            eb = torch.nn.functional.embedding_bag(torch.tensor(vals), tt_, torch.tensor(offs[:-1]), mode=mode)
            np.testing.assert_allclose(out, eb.detach().numpy(), rtol=1e-12, atol=1e-15)
            up = rng.normal(size=out.shape)
            eb.backward(torch.tensor(up))
            ids, rows = oracle.embedding_bag_backward(vals, offs, up, mode)
            dense = np.zeros_like(T); np.add.at(dense, ids, rows)
            np.testing.assert_allclose(dense, tt_.grad.numpy(), rtol=1e-12, atol=1e-15)

    def test_mlp_backward_matches_autograd(self):
        rng = np.random.default_rng(5)
        x = rng.normal(size=(9, 4)); ks = [rng.normal(size=(4, 6)), rng.normal(size=(6, 3))]
        bs = [rng.normal(size=6), rng.normal(size=3)]
        out, acts = oracle.mlp_forward(x, ks, bs)
        up = rng.normal(size=out.shape)
        dx, dk, db = oracle.mlp_backward(up, ks, acts)
        tx = torch.tensor(x, requires_grad=True); tk = [torch.tensor(k, requires_grad=True) for k in ks]
        tb = [torch.tensor(b, requires_grad=True) for b in bs]
        y = torch.relu(tx @ tk[0] + tb[0]) @ tk[1] + tb[1]
        y.backward(torch.tensor(up))
        np.testing.assert_allclose(dx, tx.grad.numpy(), rtol=1e-12)
        for l in range(2):
            np.testing.assert_allclose(dk[l], tk[l].grad.numpy(), rtol=1e-12)
            np.testing.assert_allclose(db[l], tb[l].grad.numpy(), rtol=1e-12)

    def test_dedup_is_first_occurrence_order(self):
        ids = np.array([7, 3, 7, 9, 3, 3])
        rows = np.arange(12, dtype=np.float64).reshape(6, 2)
        u, s, first = oracle.dedup_sparse_grad(ids, rows)
        assert u.tolist() == [7, 3, 9] and first.tolist() == [0, 1, 3]
        np.testing.assert_array_equal(s, [rows[0] + rows[2], rows[1] + rows[4] + rows[5], rows[3]])

    def test_adagrad_matches_torch(self):
        rng = np.random.default_rng(11)
        w = rng.normal(size=(5, 3)); g = rng.normal(size=(5, 3))
        tw = torch.tensor(w.copy(), requires_grad=True)
        opt = torch.optim.Adagrad([tw], lr=0.1, initial_accumulator_value=0.1, eps=0.0)
        tw.grad = torch.tensor(g); opt.step()
        w1, _ = oracle.adagrad_dense(w, np.full_like(w, 0.1), g, lr=0.1, eps=0.0)
        np.testing.assert_allclose(w1, tw.detach().numpy(), rtol=1e-12)

    def test_keras_sparse_adam_touches_every_row_lazy_does_not(self):
        rng = np.random.default_rng(2)
        T = rng.normal(size=(6, 2)); m = rng.normal(size=(6, 2)) * 0.1; v = np.abs(rng.normal(size=(6, 2))) * 0.1
        ids = np.array([1, 1, 4]); rows = rng.normal(size=(3, 2))
        t_k, _, _ = oracle.adam_sparse_keras(T, m, v, ids, rows, step=3)
        t_l, _, _, u = oracle.lazy_adam_sparse(T, m, v, ids, rows, step=3)
        assert (t_k[0] != T[0]).all()            # untouched row still moves under Keras Adam
        assert (t_l[0] == T[0]).all() and u.tolist() == [1, 4]
        np.testing.assert_allclose(t_k[[1, 4]], t_l[[1, 4]], rtol=1e-12)


    def test_lazy_adam_matches_torch_sparse_adam_over_three_steps(self):
        """torch.optim.SparseAdam is the same lazy rule written independently (moments of the rows present in the gradient
        only, duplicates coalesced first, step size lr * sqrt(1 - b2^t) / (1 - b1^t), eps outside the square root).  The
        oracle forms 1 - beta in fp32 as Keras does (0.00100005 for beta2): 5e-5 relative on v, hence the tolerance."""
        rng = np.random.default_rng(21)
        V, d = 12, 4
        w = rng.normal(size=(V, d))
        w0, seen = w.copy(), set()
        tw = torch.nn.Parameter(torch.tensor(w.copy()))
        opt = torch.optim.SparseAdam([tw], lr=0.01, betas=(0.9, 0.999), eps=1e-7)
        m, v = np.zeros_like(w), np.zeros_like(w)
        for step in range(1, 4):
            ids = rng.integers(0, V, size=7)
            seen.update(ids.tolist())
            rows = rng.normal(size=(7, d))
            tw.grad = torch.sparse_coo_tensor(torch.tensor(ids)[None, :], torch.tensor(rows), size=(V, d))
            opt.step()
            w, m, v, u = oracle.lazy_adam_sparse(w, m, v, ids, rows, step=step, lr=0.01)
            assert u.tolist() == list(dict.fromkeys(ids.tolist()))                      # tf.unique order
            np.testing.assert_allclose(w, tw.detach().numpy(), rtol=2e-4, atol=1e-7)
        never = np.setdiff1d(np.arange(V), sorted(seen))
        assert np.array_equal(w[never], w0[never]) and np.array_equal(m[never], np.zeros((len(never), d)))


class TestTopK:
    def test_ties_resolve_to_lower_index(self):
        s = np.array([[1.0, 3.0, 3.0, 2.0, 3.0]])
        v, i = oracle.top_k(s, 4)
        assert i.tolist() == [[1, 2, 4, 3]] and v.tolist() == [[3, 3, 3, 2]]

    def test_blocked_equals_unblocked_with_ties(self):
        rng = synth.rng_for(1)
        q = synth.exact_matrix(rng, 9, 16, 2); c = synth.exact_matrix(rng, 300, 16, 2)
        s_ref, i_ref = oracle.top_k(q.astype(np.float64) @ c.astype(np.float64).T, 20)
        s, i = oracle.brute_force_topk(q, c, 20, block=64)
        assert np.array_equal(i, i_ref) and np.array_equal(s, s_ref)
        ident = np.arange(300)[::-1].copy()
        _, ii = oracle.brute_force_topk(q, c, 20, identifiers=ident, block=128)
        assert np.array_equal(ii, ident[i_ref])

    def test_merge_of_shards_equals_global(self):
        rng = synth.rng_for(2)
        q = synth.exact_matrix(rng, 5, 8, 2); c = synth.exact_matrix(rng, 200, 8, 2)
        s_ref, i_ref = oracle.brute_force_topk(q, c, 10)
        parts = [oracle.brute_force_topk(q, c[lo:lo + 50], 10, identifiers=np.arange(lo, lo + 50)) for lo in range(0, 200, 50)]
        s, i = oracle.topk_merge([p[0] for p in parts], [p[1] for p in parts], 10)
        assert np.array_equal(i, i_ref) and np.array_equal(s, s_ref)

    def test_factorized_topk_score_and_id_mode(self):
        rng = np.random.default_rng(9)
        cands = rng.normal(size=(40, 4)); q = rng.normal(size=(6, 4))
        true_idx = np.array([0, 5, 9, 11, 30, 39])
        m = oracle.FactorizedTopKOracle(cands, ks=(1, 5, 40))
        m.update_state(q, cands[true_idx])
        res = m.result()
        scores = q @ cands.T
        rank = (scores > scores[np.arange(6), true_idx][:, None]).sum(1)
        for k in (1, 5, 40):
            assert res[f"factorized_top_k/top_{k}_categorical_accuracy"] == pytest.approx((rank < k).mean())
        m2 = oracle.FactorizedTopKOracle(cands, ks=(1, 5, 40))
        m2.update_state(q, cands[true_idx], true_candidate_ids=true_idx)
        assert m2.result() == res                                  # tie-free data: both modes agree


class TestTrainStep:
    def test_two_tower_step_decreases_loss_and_updates_only_touched_rows(self):
        rng = synth.rng_for(3)
        qs = oracle.TowerSpec([("user", "id", 50, None)], 8, (16, 8))
        cs = oracle.TowerSpec([("item", "id", 40, None), ("cat", "bag", 10, "mean")], 8, (16, 8))
        qp, cp = oracle.init_tower(qs, rng, np.float64), oracle.init_tower(cs, rng, np.float64)
        mk = lambda p: {"tables": {k: np.full_like(v, 0.1) for k, v in p["tables"].items()},
                        "kernels": [np.full_like(k, 0.1) for k in p["kernels"]],
                        "biases": [np.full_like(b, 0.1) for b in p["biases"]]}
        qsl, csl = mk(qp), mk(cp)
        bq = {"user": synth.draw_ids(rng, 12, 50)}
        bc = {"item": synth.draw_ids(rng, 12, 40), "cat": synth.draw_bags(rng, 12, 10, 0, 3)}
        before = qp["tables"]["user"].copy()
        losses = []
        for _ in range(5):
            r = oracle.two_tower_train_step(qs, cs, qp, cp, qsl, csl, bq, bc, temperature=0.5, lr=0.1, l2=1e-3)
            losses.append(r["total_loss"])
        assert losses[-1] < losses[0]
        untouched = np.setdiff1d(np.arange(50), bq["user"])
        assert np.array_equal(qp["tables"]["user"][untouched], before[untouched])
        assert r["regularization_loss"] > 0

    def test_full_train_step_matches_an_independent_torch_autograd_model(self):
        """End to end against an INDEPENDENT statement of the same published semantics: torch modules (embedding lookups,
        mean-pooled bag with duplicates and an empty bag, Dense(relu) -> Dense, logits / T, cross entropy with
        reduction = sum against the diagonal, L2 kernel regulariser), autograd for every gradient, and Keras Adagrad written
        out (acc += g^2; w -= lr * g / sqrt(acc + eps); on a table, g = 0 leaves a row and its accumulator alone).  Two
        steps, so that the second one runs on updated accumulators."""
        rng = synth.rng_for(8)
        d, mlp, T, lr, l2, B = 8, (16, 8), 0.25, 0.1, 1e-3, 14
        qs = oracle.TowerSpec([("user", "id", 30, None)], d, mlp)
        cs = oracle.TowerSpec([("item", "id", 25, None), ("cat", "bag", 9, "mean")], d, mlp)
        qp, cp = oracle.init_tower(qs, rng, np.float64), oracle.init_tower(cs, rng, np.float64)
        mk = lambda p: {"tables": {k: np.full_like(v, 0.1) for k, v in p["tables"].items()},
                        "kernels": [np.full_like(k, 0.1) for k in p["kernels"]],
                        "biases": [np.full_like(b, 0.1) for b in p["biases"]]}
        qsl, csl = mk(qp), mk(cp)
        t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), requires_grad=True)
        P = {"user": t(qp["tables"]["user"]), "item": t(cp["tables"]["item"]), "cat": t(cp["tables"]["cat"]),
             "qk": [t(k) for k in qp["kernels"]], "qb": [t(b) for b in qp["biases"]],
             "ck": [t(k) for k in cp["kernels"]], "cb": [t(b) for b in cp["biases"]]}
        leaves = [P["user"], P["item"], P["cat"], *P["qk"], *P["qb"], *P["ck"], *P["cb"]]
        acc = [torch.full_like(v, 0.1) for v in leaves]
        for step in range(2):
            users, items = synth.draw_ids(rng, B, 30, 1.2), synth.draw_ids(rng, B, 25, 1.2)
            vals, offs = synth.draw_bags(rng, B, 9, 1, 3, empty_frac=0.25)
            assert (np.diff(offs) == 0).any() and len(np.unique(vals)) < len(vals)      # empty bags and duplicates present
            r = oracle.two_tower_train_step(qs, cs, qp, cp, qsl, csl, {"user": users}, {"item": items, "cat": (vals, offs)},
                                            temperature=T, lr=lr, l2=l2)
            lens = np.diff(offs)
            bag_of = torch.tensor(np.repeat(np.arange(B), lens))
            pooled = torch.zeros((B, d), dtype=torch.float64).index_add(0, bag_of, P["cat"][torch.tensor(vals)])
            pooled = pooled / torch.tensor(np.maximum(lens, 1), dtype=torch.float64)[:, None]
            mlp_t = lambda x, ks, bs: torch.relu(x @ ks[0] + bs[0]) @ ks[1] + bs[1]
            q = mlp_t(P["user"][torch.tensor(users)], P["qk"], P["qb"])
            c = mlp_t(P["item"][torch.tensor(items)] + pooled, P["ck"], P["cb"])
            loss = torch.nn.functional.cross_entropy(q @ c.T / T, torch.arange(B), reduction="sum")
            reg = l2 * sum((k * k).sum() for k in P["qk"] + P["ck"])
            assert r["loss"] == pytest.approx(float(loss.detach()), rel=1e-12)
            assert r["regularization_loss"] == pytest.approx(float(reg.detach()), rel=1e-12)
            grads = torch.autograd.grad(loss + reg, leaves)
            with torch.no_grad():
                for v, a, g in zip(leaves, acc, grads):
                    a += g * g
                    v -= lr * g / torch.sqrt(a + 1e-7)
            got = [qp["tables"]["user"], cp["tables"]["item"], cp["tables"]["cat"], *qp["kernels"], *qp["biases"],
                   *cp["kernels"], *cp["biases"]]
            slots = [qsl["tables"]["user"], csl["tables"]["item"], csl["tables"]["cat"], *qsl["kernels"], *qsl["biases"],
                     *csl["kernels"], *csl["biases"]]
            for v, a, w, sa in zip(leaves, acc, got, slots):
                np.testing.assert_allclose(w, v.detach().numpy(), rtol=1e-10, atol=1e-13)
                np.testing.assert_allclose(sa, a.numpy(), rtol=1e-10, atol=1e-13)
            assert np.array_equal(np.sort(r["unique"]["c/cat"]), np.unique(vals))

    def test_bf16_round_is_rne(self):
        x = np.array([1.0, 1.00390625, 1.01171875, -2.5, 3.14159], dtype=np.float32)
        ref = torch.tensor(x).to(torch.bfloat16).float().numpy()
        assert np.array_equal(oracle.bf16_round(x), ref)
