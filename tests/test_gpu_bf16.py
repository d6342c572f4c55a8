"""GPU parity, bf16 precision (tcgen05 tensor-core kernels) through the C-ABI against the
oracle.  Tolerance 2e-2 relative (north_star bf16 bound); on exact-arithmetic inputs (small
dyadic values: every product and partial sum representable) the fp32-accumulated results must
be EXACT, which pins descriptor/layout bugs that a loose tolerance would hide."""
import numpy as np
import pytest
import torch

import oracle
from two_tower_b200 import synth

pytestmark = pytest.mark.gpu
BF16_RTOL = 2e-2


def dev(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def bf(x):
    return dev(x, torch.bfloat16)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="module")
def ops(tt):
    tt.ops.device_check()
    return tt.ops


class TestDenseTensorCore:
    @pytest.mark.parametrize("M,i,o", [(128, 64, 128), (256, 128, 256), (8192, 128, 256), (8192, 256, 128), (1000, 64, 48), (72, 320, 16)])
    def test_forward_exact_on_dyadic_inputs(self, ops, M, i, o):
        rng = synth.rng_for(M + i + o)
        x = synth.exact_matrix(rng, M, i, 4); k = synth.exact_matrix(rng, i, o, 4)
        b = synth.exact_matrix(rng, 1, o, 4)[0]
        y, y_f = ops.dense_fwd("bf16", bf(x), bf(k), dev(b), relu=False, want_f32=True)
        ref = x.astype(np.float64) @ k + b
        assert np.array_equal(y_f.cpu().numpy().astype(np.float64), ref)
        assert np.array_equal(y.float().cpu().numpy(), oracle.bf16_round(ref.astype(np.float32)))
        y2, _ = ops.dense_fwd("bf16", bf(x), bf(k), dev(b), relu=True)
        assert np.array_equal(y2.float().cpu().numpy(), oracle.bf16_round(np.maximum(ref, 0).astype(np.float32)))

    @pytest.mark.parametrize("M,i,o", [(256, 128, 256), (8192, 128, 256), (8192, 256, 128), (1000, 64, 48)])
    def test_backward_exact_on_dyadic_inputs(self, ops, M, i, o):
        rng = synth.rng_for(M * 3 + i + o)
        x = np.maximum(synth.exact_matrix(rng, M, i, 4), 0); k = synth.exact_matrix(rng, i, o, 2)
        dy = synth.exact_matrix(rng, M, o, 2)
        dx, dx_f, dk, P, db = ops.dense_bwd("bf16", bf(dy), bf(x), bf(k), relu_mask_x=True, want_dx=True, want_dx_f32=True)
        ref_dx = (dy.astype(np.float64) @ k.T) * (x > 0)
        assert np.array_equal(dx_f.cpu().numpy().astype(np.float64), ref_dx)
        assert np.array_equal(dx.float().cpu().numpy(), oracle.bf16_round(ref_dx.astype(np.float32)))
        assert P == dk.shape[0] == db.shape[0] and P >= 1
        assert np.array_equal(dk.sum(0).cpu().numpy().astype(np.float64), x.astype(np.float64).T @ dy)
        assert np.array_equal(db.sum(0).cpu().numpy().astype(np.float64), dy.astype(np.float64).sum(0))

    def test_gaussian_within_bf16_tolerance(self, ops):
        rng = synth.rng_for(12)
        M, i, o = 4096, 128, 256
        x = rng.normal(size=(M, i)).astype(np.float32); k = oracle.glorot_uniform(rng, i, o); b = rng.normal(size=o).astype(np.float32)
        y, _ = ops.dense_fwd("bf16", bf(x), bf(k), dev(b), relu=True)
        ref = oracle.dense_forward(x.astype(np.float64), k, b, "relu")
        assert rel_err(y.float().cpu().numpy(), ref) < BF16_RTOL


class TestRetrievalTensorCore:
    def _inputs(self, nq, nc, d, seed, scale=0.3):
        rng = synth.rng_for(seed)
        q = oracle.bf16_round(rng.normal(size=(nq, d)).astype(np.float32) * scale)
        c = oracle.bf16_round(rng.normal(size=(nc, d)).astype(np.float32) * scale)
        return rng, q, c

    def _check(self, ops, q, c, label_offset=0, **kw):
        T = kw.get("temperature")
        inv_t = 1.0 if T is None else 1.0 / T
        w, p, ids = kw.get("sample_weight"), kw.get("candidate_sampling_probability"), kw.get("candidate_ids")
        logq = None if p is None else dev(np.log(np.clip(p, 1e-6, 1.0)), torch.float32)
        ids_d = dev(ids) if kw.get("remove_accidental_hits") else None
        w_d = None if w is None else dev(w, torch.float32)
        nq = q.shape[0]
        perm = np.concatenate([np.arange(label_offset, label_offset + nq), np.arange(0, label_offset),
                               np.arange(label_offset + nq, c.shape[0])])
        kw_o = dict(kw)
        if p is not None: kw_o["candidate_sampling_probability"] = np.asarray(p)[perm]
        if ids is not None: kw_o["candidate_ids"] = np.asarray(ids)[perm]
        r = oracle.retrieval_loss_and_grads(q.astype(np.float64), c.astype(np.float64)[perm], **kw_o)
        qb, cb = bf(q), bf(c)
        loss, lse, pos = ops.retrieval_loss_fwd("bf16", qb, cb, inv_t, label_offset, w_d, logq, ids_d)
        # inputs are bf16-exact, so the forward differs from fp64 only by fp32 accumulation + ex2.approx
        assert float(loss.item()) == pytest.approx(r["loss"], rel=2e-4)
        assert rel_err(lse.cpu().numpy(), r["lse"]) < 2e-4 and rel_err(pos.cpu().numpy(), r["pos"]) < 2e-4
        g = ops.retrieval_loss_bwd("bf16", qb, cb, inv_t, lse, label_offset, w_d, logq, ids_d, want_bf16=(True, True))
        dc_ref = np.empty_like(r["dc"]); dc_ref[perm] = r["dc"]
        assert rel_err(g["dq"].cpu().numpy(), r["dq"]) < BF16_RTOL
        assert rel_err(g["dc"].cpu().numpy(), dc_ref) < BF16_RTOL
        assert np.array_equal(g["dq_bf16"].float().cpu().numpy(), oracle.bf16_round(g["dq"].cpu().numpy()))
        assert np.array_equal(g["dc_bf16"].float().cpu().numpy(), oracle.bf16_round(g["dc"].cpu().numpy()))
        # forward + dQ in one pass (flash-attention forward shape) and the dC-only backward that goes with it
        if logq is None and ids_d is None and ops.retrieval_fwd_dq_supported(nq, c.shape[0], q.shape[1]):
            loss2, lse2, pos2, dq2, fws = ops.retrieval_loss_fwd_dq(qb, cb, inv_t, label_offset, w_d, fork=True)
            ops.join_side_work()
            if w_d is None:                                   # dC pass fed by the fold's -lse2 array (broadcast loads)
                dc_f = ops.retrieval_loss_bwd_dc_fused(qb, cb, inv_t, fws, label_offset, None, 1.0)
                assert rel_err(dc_f.sum(0).cpu().numpy(), dc_ref) < BF16_RTOL
            assert float(loss2.item()) == pytest.approx(r["loss"], rel=2e-4)
            assert rel_err(lse2.cpu().numpy(), r["lse"]) < 2e-4 and rel_err(pos2.cpu().numpy(), r["pos"]) < 2e-4
            assert rel_err(dq2.cpu().numpy(), r["dq"]) < BF16_RTOL
            none, dc_parts = ops.retrieval_loss_bwd_parts(qb, cb, inv_t, lse2, label_offset, w_d, None, None, 1.0, want_dq=False)
            assert none is None
            assert rel_err(dc_parts.sum(0).cpu().numpy(), dc_ref) < BF16_RTOL
        return r

    @pytest.mark.parametrize("nq,nc,d,off", [(128, 128, 64, 0), (256, 256, 128, 0), (1000, 1000, 128, 0), (8, 8, 64, 0), (77, 203, 64, 100),
                                             (520, 1304, 256, 264), (384, 2048, 192, 1024), (4096, 4096, 128, 0)])
    def test_shapes_splits_and_label_offset(self, ops, nq, nc, d, off):
        _, q, c = self._inputs(nq, nc, d, nq + nc + d)
        self._check(ops, q, c, off, temperature=0.25)

    def test_fused_forward_dq_rescales_when_the_row_maximum_keeps_rising(self, ops):
        """Candidates ordered by increasing norm: the row maxima grow from tile to tile by far more than the lazy
        threshold (8 in the log2 domain), so the accumulator rescale of the one-pass forward + dQ kernel runs."""
        rng, q, c = self._inputs(300, 2304, 128, 11, scale=0.5)
        c = oracle.bf16_round(c * np.linspace(0.05, 3.0, c.shape[0], dtype=np.float32)[:, None])
        self._check(ops, q, c, 1000, temperature=0.1)
        self._check(ops, q, c[::-1].copy(), 0, temperature=0.1)

    def test_no_temperature(self, ops):
        _, q, c = self._inputs(512, 512, 128, 3, scale=0.5)
        self._check(ops, q, c)

    def test_weights_logq_and_accidental_hits(self, ops):
        rng, q, c = self._inputs(640, 1152, 128, 5)
        ids = synth.draw_ids(rng, 1152, 300, zipf=1.2)          # many duplicate candidate ids
        self._check(ops, q, c, 256, temperature=0.2, sample_weight=rng.uniform(0.5, 1.5, 640),
                    candidate_sampling_probability=rng.uniform(1e-4, 0.3, 1152), candidate_ids=ids,
                    remove_accidental_hits=True)
        self._check(ops, q, c, 0, temperature=0.2, sample_weight=rng.uniform(0.5, 1.5, 640))

    def test_golden_cfg1_within_bf16_tolerance(self, ops, golden_dir):
        g = np.load(golden_dir / "cfg1.npz")
        cfg = synth.CONFIGS["cfg1"]; rng = synth.rng_for(cfg.seed)
        U = oracle.keras_uniform(rng, (cfg.v_user, cfg.dim)); I = oracle.keras_uniform(rng, (cfg.v_item, cfg.dim))
        qb, cb = bf(U[g["uid"]]), bf(I[g["iid"]])
        loss, lse, _ = ops.retrieval_loss_fwd("bf16", qb, cb, 1.0 / cfg.temperature)
        assert float(loss.item()) == pytest.approx(float(g["temp_loss"]), rel=BF16_RTOL)
        r = ops.retrieval_loss_bwd("bf16", qb, cb, 1.0 / cfg.temperature, lse)
        assert rel_err(r["dq"].cpu().numpy(), g["temp_dq"]) < BF16_RTOL
        assert rel_err(r["dc"].cpu().numpy(), g["temp_dc"]) < BF16_RTOL

    def test_full_size_uniform_embeddings_property(self, ops):
        B, d = 8192, 128
        q = torch.full((B, d), 0.0625, device="cuda", dtype=torch.bfloat16); c = q.clone()
        loss, lse, _ = ops.retrieval_loss_fwd("bf16", q, c, 10.0)
        assert float(loss.item()) == pytest.approx(B * np.log(B), rel=1e-4)
        g = ops.retrieval_loss_bwd("bf16", q, c, 10.0, lse)
        assert float(g["dq"].abs().max().item()) < 1e-3 and float(g["dc"].abs().max().item()) < 1e-3


    def test_one_workspace_across_shapes(self, ops):
        """A caller re-uses one workspace for every shape: the partials a small shape leaves behind must never be read as
        arrival tickets by a larger one (round 2: the ticket area has one size for all shapes)."""
        B, d = 8192, 128
        q = torch.full((B, d), 0.0625, device="cuda", dtype=torch.bfloat16); c = q.clone()
        loss0, lse0, _ = ops.retrieval_loss_fwd("bf16", q, c, 10.0)            # sizes the cached workspace for the big shape
        rng = synth.rng_for(5)
        for nq, nc in ((128, 256), (200, 1000), (1024, 1024)):
            qs = bf((rng.normal(size=(nq, d)) * 0.1).astype(np.float32)); cs = bf((rng.normal(size=(nc, d)) * 0.1).astype(np.float32))
            _l, lse, _p = ops.retrieval_loss_fwd("bf16", qs, cs, 10.0)
            ops.retrieval_loss_bwd("bf16", qs, cs, 10.0, lse)                     # dQ / dC partials live in the same workspace
        loss1, lse1, _ = ops.retrieval_loss_fwd("bf16", q, c, 10.0)
        assert float(loss1.item()) == float(loss0.item()) == pytest.approx(B * np.log(B), rel=1e-4)
        assert torch.equal(lse0, lse1)


class TestTrainStepBf16:
    def _model(self, tt, vu, vi, d, mlp, T, lr):
        tt.set_precision("bf16")
        # layer initialisers are seeded by (config.seed, process-wide layer counter): pin both so that the statistical
        # bounds of these tests do not depend on which tests built layers before them
        tt.set_seed(21)
        tt.layers._layer_counter[0] = 0

        class TwoTower(tt.models.Model):
            def __init__(s):
                super().__init__()
                s.user_model = tt.Sequential([tt.layers.Embedding(vu, d), tt.layers.Dense(mlp[0], "relu"), tt.layers.Dense(mlp[1])])
                s.item_model = tt.Sequential([tt.layers.Embedding(vi, d), tt.layers.Dense(mlp[0], "relu"), tt.layers.Dense(mlp[1])])
                s.task = tt.tasks.Retrieval(temperature=T)

            def compute_loss(s, f, training=False):
                s.last_q = s.user_model(f["user_id_encoded"]); s.last_c = s.item_model(f["item_id_encoded"])
                return s.task(s.last_q, s.last_c)

        model = TwoTower(); model.compile(optimizer=tt.optimizers.Adagrad(lr))
        return model

    def test_loss_embeddings_and_gradients_within_bf16_tolerance(self, tt):
        vu, vi, d, mlp, B, T, lr = 2000, 1500, 64, (128, 64), 512, 0.5, 0.05
        model = self._model(tt, vu, vi, d, mlp, T, lr)
        rng = synth.rng_for(31)
        mkb = lambda: {"user_id_encoded": synth.draw_ids(rng, B, vu, 1.3), "item_id_encoded": synth.draw_ids(rng, B, vi, 1.3)}
        model.test_step(mkb())
        qs = oracle.TowerSpec([("user_id_encoded", "id", vu, None)], d, mlp)
        cs = oracle.TowerSpec([("item_id_encoded", "id", vi, None)], d, mlp)
        get = lambda seq, name: {"tables": {name: seq.layers[0].get_weights()[0].astype(np.float64)},
                                 "kernels": [l.get_weights()[0].astype(np.float64) for l in seq.layers[1:]],
                                 "biases": [l.get_weights()[1].astype(np.float64) for l in seq.layers[1:]]}
        qp, cp = get(model.user_model, "user_id_encoded"), get(model.item_model, "item_id_encoded")
        b = mkb()
        # gradients of one step, taken from the tape exactly as train_step does
        with tt.GradientTape() as tape:
            loss = model.compute_loss(b, training=True)
            variables = model.trainable_variables
            grads = tape.gradient(loss, variables)
        # bf16=True: the oracle rounds the inputs of every matmul to bf16 like the tensor-core path, so
        # ReLU masks are decided on the same rounded pre-activations (a mask flip is a discontinuity no
        # tolerance covers); accumulation stays fp64
        q_ref, qc = oracle.tower_forward(qs, qp, {"user_id_encoded": b["user_id_encoded"]}, bf16=True)
        c_ref, cc = oracle.tower_forward(cs, cp, {"item_id_encoded": b["item_id_encoded"]}, bf16=True)
        q_ref, c_ref = oracle.bf16_round(q_ref).astype(np.float64), oracle.bf16_round(c_ref).astype(np.float64)
        r = oracle.retrieval_loss_and_grads(q_ref, c_ref, temperature=T)
        assert loss.item() == pytest.approx(r["loss"], rel=BF16_RTOL)
        assert rel_err(model.last_q.numpy(), q_ref) < BF16_RTOL and rel_err(model.last_c.numpy(), c_ref) < BF16_RTOL
        dk, db, sparse = oracle.tower_backward(qs, qp, {"user_id_encoded": b["user_id_encoded"]}, qc, r["dq"])
        by_name = {v.name: g for v, g in zip(variables, grads)}
        emb = by_name[model.user_model.layers[0].embeddings.name]
        assert np.array_equal(emb.values.cpu().numpy(), b["user_id_encoded"])            # lookup indices bit-exact
        assert rel_err(emb.rows.cpu().numpy(), sparse["user_id_encoded"][1]) < BF16_RTOL
        for j, layer in enumerate(model.user_model.layers[1:]):
            gk = by_name[layer.kernel.name]; gb = by_name[layer.bias.name]
            # weight gradients are cancelling sums over the batch of bf16-rounded terms: north_star
            # states no tolerance for them; measured 1-2 % of max |grad|, bounded here at 5 %
            assert rel_err(gk.parts.sum(0).cpu().numpy(), dk[j]) < 5e-2
        # bias gradients are column sums of dy that cancel almost completely (for the last layer
        # sum_i (softmax - eye)_ij ~ 0), so they are bounded by the rounding of their terms instead:
        # |err_j| <= 2^-7 * sum_i |dy_ij|
        dys = [None, r["dq"]]
        g1 = r["dq"] @ qp["kernels"][1].T * (qc["acts"][1] > 0)
        dys[0] = g1
        for j, layer in enumerate(model.user_model.layers[1:]):
            gb = by_name[layer.bias.name].parts.sum(0).cpu().numpy().reshape(-1)
            assert (np.abs(gb - db[j]) <= 2.0 ** -7 * np.abs(dys[j]).sum(0) + 1e-6).all()

    def test_three_steps_track_the_oracle(self, tt):
        vu, vi, d, mlp, B, T, lr = 2000, 1500, 64, (128, 64), 512, 0.5, 0.05
        model = self._model(tt, vu, vi, d, mlp, T, lr)
        rng = synth.rng_for(32)
        mkb = lambda: {"user_id_encoded": synth.draw_ids(rng, B, vu, 1.3), "item_id_encoded": synth.draw_ids(rng, B, vi, 1.3)}
        model.test_step(mkb())
        qs = oracle.TowerSpec([("user_id_encoded", "id", vu, None)], d, mlp)
        cs = oracle.TowerSpec([("item_id_encoded", "id", vi, None)], d, mlp)
        get = lambda seq, name: {"tables": {name: seq.layers[0].get_weights()[0].astype(np.float64)},
                                 "kernels": [l.get_weights()[0].astype(np.float64) for l in seq.layers[1:]],
                                 "biases": [l.get_weights()[1].astype(np.float64) for l in seq.layers[1:]]}
        qp, cp = get(model.user_model, "user_id_encoded"), get(model.item_model, "item_id_encoded")
        tab0 = qp["tables"]["user_id_encoded"].copy()
        mk = lambda p: {"tables": {k: np.full(v.shape, 0.1) for k, v in p["tables"].items()},
                        "kernels": [np.full(k.shape, 0.1) for k in p["kernels"]], "biases": [np.full(x.shape, 0.1) for x in p["biases"]]}
        qsl, csl = mk(qp), mk(cp)
        touched = set()
        for _ in range(3):
            b = mkb()
            out = model.train_step(b)
            ref = oracle.two_tower_train_step(qs, cs, qp, cp, qsl, csl, {"user_id_encoded": b["user_id_encoded"]},
                                              {"item_id_encoded": b["item_id_encoded"]}, temperature=T, lr=lr, bf16=True)
            assert float(out["loss"].item()) == pytest.approx(ref["loss"], rel=BF16_RTOL)
            touched |= set(ref["unique"]["q/user_id_encoded"].tolist())
        tab = model.user_model.layers[0].get_weights()[0].astype(np.float64)
        rows = np.array(sorted(touched)); rest = np.setdiff1d(np.arange(vu), rows)
        assert np.array_equal(tab[rest], tab0[rest])                       # gradient row set exact: nothing else moved
        assert (np.abs(tab[rows] - tab0[rows]).max(axis=1) > 0).all()      # every touched row moved
        du, dr = (tab - tab0)[rows].ravel(), (qp["tables"]["user_id_encoded"] - tab0)[rows].ravel()
        # Adagrad's g/sqrt(acc+g^2) amplifies bf16 gradient noise on near-zero components, so the
        # updated table is compared by direction/magnitude, not element-wise at 2e-2
        assert float(du @ dr / (np.linalg.norm(du) * np.linalg.norm(dr))) > 0.98
        assert np.linalg.norm(du) == pytest.approx(np.linalg.norm(dr), rel=5e-2)


class TestTopKTensorCore:
    def _check(self, ops, q, c, k, identifiers=None, base=0):
        s, i = ops.topk_bruteforce("bf16", bf(q), bf(c), k, cand_index_base=base,
                                   identifiers=None if identifiers is None else dev(identifiers))
        ref_s, ref_i = oracle.brute_force_topk(q, c, k, identifiers)
        assert np.array_equal(i.cpu().numpy() - (0 if identifiers is not None else base), ref_i)
        assert np.array_equal(s.cpu().numpy().astype(np.float64), ref_s)

    @pytest.mark.parametrize("nq,nc,d,k", [(128, 1024, 128, 100), (70, 1000, 64, 100), (5, 37, 64, 37), (300, 5000, 128, 10),
                                           (64, 3000, 256, 64), (257, 4099, 192, 1), (1000, 20000, 128, 100)])
    def test_exact_arithmetic_data_bit_exact_ids_with_ties(self, ops, nq, nc, d, k):
        rng = synth.rng_for(nq + nc + k)
        q, c = synth.exact_matrix(rng, nq, d, 2), synth.exact_matrix(rng, nc, d, 2)
        self._check(ops, q, c, k)

    def test_split_candidates_identifiers_and_base(self, ops, tt):
        rng = synth.rng_for(199)
        nq, nc, d, k = 40, 60000, 64, 100
        assert tt._lib.load().tt_topk_num_splits(1, nq, nc, d, k) > 1
        q, c = synth.exact_matrix(rng, nq, d, 3), synth.exact_matrix(rng, nc, d, 3)
        ident = rng.permutation(nc).astype(np.int64) + 10_000_000_000
        self._check(ops, q, c, k, identifiers=ident)
        self._check(ops, q, c, k, base=777)

    @pytest.mark.parametrize("nq,nc,d,k", [(512, 50000, 128, 100), (300, 20000, 64, 10), (4096, 200000, 128, 100),
                                           (64, 100000, 256, 64)])
    def test_gaussian_bf16_bit_exact_ids(self, ops, nq, nc, d, k):
        """Real-valued data: the ids equal tf.math.top_k on the correctly rounded fp32 score tensor (exact re-rank of
        the k + 16 pool, csrc/topk_rerank.cu) -- bit-exact, not 'up to near-tie swaps'."""
        rng = synth.rng_for(17 + nq)
        q = oracle.bf16_round((rng.normal(size=(nq, d)) / np.sqrt(d)).astype(np.float32))
        c = oracle.bf16_round((rng.normal(size=(nc, d)) / np.sqrt(d)).astype(np.float32))
        unc = torch.zeros(1, dtype=torch.int32, device="cuda")
        s, i = ops.topk_bruteforce("bf16", bf(q), bf(c), k, uncertain=unc)
        ref_s, ref_i = oracle.brute_force_topk(q, c, k, score_dtype=np.float32, block=16384)
        assert np.array_equal(i.cpu().numpy(), ref_i)
        assert np.array_equal(s.cpu().numpy(), ref_s)
        assert int(unc.item()) == 0

    def test_cfg5_distribution_10m_candidates_bit_exact_ids(self, ops):
        """BASELINE configs[4] at full candidate count: 10 M x 128 bf16 candidates ~ N(0,1)/sqrt(d), top-100, a
        256-query subset (the oracle's fp64 scoring of 2.56 G pairs takes ~1 min on the host)."""
        nq, nc, d, k = 256, 10_000_000, 128, 100
        g = torch.Generator(device="cuda"); g.manual_seed(5678)
        cand = (torch.randn((nc, d), device="cuda", generator=g) / d ** 0.5).to(torch.bfloat16)
        q = (torch.randn((nq, d), device="cuda", generator=g) / d ** 0.5).to(torch.bfloat16)
        unc = torch.zeros(1, dtype=torch.int32, device="cuda")
        s, i = ops.topk_bruteforce("bf16", q, cand, k, uncertain=unc)
        c_host = cand.float().cpu().numpy(); q_host = q.float().cpu().numpy()
        ref_s, ref_i = oracle.brute_force_topk(q_host, c_host, k, score_dtype=np.float32, block=65536)
        assert np.array_equal(i.cpu().numpy(), ref_i)
        assert np.array_equal(s.cpu().numpy(), ref_s)
        assert int(unc.item()) == 0

    def test_brute_force_layer_and_metric(self, tt):
        tt.set_precision("bf16")
        rng = synth.rng_for(23)
        cands = synth.exact_matrix(rng, 5000, 64, 3); q = synth.exact_matrix(rng, 256, 64, 3)
        index = tt.layers.factorized_top_k.BruteForce(k=50).index(torch.as_tensor(cands).cuda(), torch.arange(5000) * 3)
        s, i = index(torch.as_tensor(q).cuda())
        ref_s, ref_i = oracle.brute_force_topk(q, cands, 50, identifiers=np.arange(5000) * 3)
        assert np.array_equal(i.cpu().numpy(), ref_i) and index.is_exact()
        excl = ref_i[:, :3].copy()
        s2, i2 = index.query_with_exclusions(torch.as_tensor(q).cuda(), excl, k=10)
        assert np.array_equal(i2.cpu().numpy(), ref_i[:, 3:13])
        true_idx = rng.integers(0, 5000, 256)
        m = tt.metrics.FactorizedTopK(torch.as_tensor(cands).cuda(), ks=(1, 5, 10, 20, 50, 100))
        o = oracle.FactorizedTopKOracle(cands, ks=(1, 5, 10, 20, 50, 100))
        m.update_state(torch.as_tensor(q).cuda(), torch.as_tensor(cands[true_idx]).cuda())
        o.update_state(q, cands[true_idx])
        assert m.result() == pytest.approx(o.result())


class TestOperandMajorness:
    """tcgen05 shared-memory descriptors: K-major and MN-major operands (both 128B-swizzled) must
    give the same exact product, so no kernel needs a transposed copy of its input."""

    @pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
    @pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 128, 128), (1000, 136, 200), (128, 256, 8192)])
    def test_gemm_all_layouts_exact(self, tt, a_mn, b_mn, M, N, K):
        rng = synth.rng_for(M + N + K)
        A = synth.exact_matrix(rng, M, K, 3); B = synth.exact_matrix(rng, N, K, 3)
        a_dev = bf(A.T.copy() if a_mn else A); b_dev = bf(B.T.copy() if b_mn else B)
        out = torch.empty((M, N), dtype=torch.float32, device="cuda")
        tt._lib.check(tt._lib.load().tt_debug_gemm_bf16(a_dev.data_ptr(), a_mn, b_dev.data_ptr(), b_mn, M, N, K,
                                                         out.data_ptr(), torch.cuda.current_stream().cuda_stream))
        assert np.array_equal(out.cpu().numpy().astype(np.float64), A.astype(np.float64) @ B.astype(np.float64).T)
