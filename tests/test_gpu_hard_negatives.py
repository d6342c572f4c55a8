"""tfrs.tasks.Retrieval(num_hard_negatives=n): selection by the brute-force top-k kernel, loss and
gradients on the gathered logits -- against the oracle's HardNegativeMining restatement (top_k of scores + labels * MAX,
ties -> lower index).  Selected index sets bit-exact; fp32 within 1e-5, bf16 within 2e-2."""
import numpy as np
import pytest
import torch

import oracle
from two_tower_b200 import synth

pytestmark = pytest.mark.gpu


def dev(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _selected_oracle(q, c, n, T):
    s = (q.astype(np.float64) @ c.astype(np.float64).T) / T
    nq, nc = s.shape
    boosted = s.copy(); boosted[np.arange(nq), np.arange(nq)] += oracle.MAX_FLOAT
    return np.argsort(-boosted, axis=1, kind="stable")[:, :min(n + 1, nc)]


@pytest.mark.parametrize("precision,nq,d,n,weights", [("fp32", 256, 64, 7, False), ("fp32", 300, 128, 50, True),
                                                      ("bf16", 512, 128, 15, False), ("bf16", 1000, 128, 100, True),
                                                      ("fp32", 40, 64, 100, False)])          # n + 1 > nc: everything is kept
def test_ops_against_the_oracle(tt, precision, nq, d, n, weights):
    ops = tt.ops
    rng = synth.rng_for(nq + d + n)
    T = 0.25
    q = rng.normal(size=(nq, d)).astype(np.float32) * 0.4
    c = rng.normal(size=(nq, d)).astype(np.float32) * 0.4
    if precision == "bf16":
        q, c = oracle.bf16_round(q), oracle.bf16_round(c)
    w = rng.uniform(0.5, 1.5, nq).astype(np.float32) if weights else None
    dt = torch.float32 if precision == "fp32" else torch.bfloat16
    qd, cd = dev(q, dt), dev(c, dt)
    sel = ops.select_hard_negatives(precision, qd, cd, n, 0)
    want = _selected_oracle(q, c, n, T)
    got = sel.cpu().numpy()
    assert np.array_equal(got[:, 0], np.arange(nq))                                  # the positive comes first
    assert np.array_equal(np.sort(got, axis=1), np.sort(want, axis=1))               # same index SET per row, bit-exact
    wd = None if w is None else dev(w)
    loss, lse, pos, scores = ops.hard_negative_loss_fwd(precision, qd, cd, sel, 1.0 / T, wd)
    r = oracle.retrieval_loss_and_grads(q.astype(np.float64), c.astype(np.float64), temperature=T, sample_weight=w,
                                        num_hard_negatives=n)
    tol = 1e-5 if precision == "fp32" else 2e-4        # bf16 inputs are exact here; only the fp32 accumulation differs
    assert float(loss.item()) == pytest.approx(r["loss"], rel=tol)
    assert rel_err(lse.cpu().numpy(), r["lse"]) < tol and rel_err(pos.cpu().numpy(), r["pos"]) < tol
    dq, dc = ops.hard_negative_loss_bwd(precision, qd, cd, sel, 1.0 / T, scores, lse, wd, 1.0)
    gtol = 1e-5 if precision == "fp32" else 2e-2
    assert rel_err(dq.cpu().numpy(), r["dq"]) < gtol and rel_err(dc.cpu().numpy(), r["dc"]) < gtol


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_retrieval_task_with_hard_negatives_trains(tt, precision):
    tt.set_precision(precision)
    tt.set_seed(8)
    tt.layers._layer_counter[0] = 0
    vu, vi, B = 700, 500, 384

    class TwoTower(tt.models.Model):
        def __init__(s):
            super().__init__()
            s.user_model = tt.Sequential([tt.layers.Embedding(vu, 128), tt.layers.Dense(256, "relu"), tt.layers.Dense(128)])
            s.item_model = tt.Sequential([tt.layers.Embedding(vi, 128), tt.layers.Dense(256, "relu"), tt.layers.Dense(128)])
            s.task = tt.tasks.Retrieval(temperature=0.5, num_hard_negatives=20)

        def compute_loss(s, f, training=False):
            s.last_q, s.last_c = s.user_model(f["user_id_encoded"]), s.item_model(f["item_id_encoded"])
            return s.task(s.last_q, s.last_c)

    model = TwoTower()
    model.compile(optimizer=tt.optimizers.Adagrad(0.002))   # small steps: the hardest-20 set is re-selected every step
    rng = synth.rng_for(3)
    batch = {"user_id_encoded": synth.draw_ids(rng, B, vu, None), "item_id_encoded": synth.draw_ids(rng, B, vi, None)}
    out0 = model.test_step(batch)
    qn, cn = model.last_q.numpy().astype(np.float64), model.last_c.numpy().astype(np.float64)
    ref = oracle.retrieval_loss(qn, cn, temperature=0.5, num_hard_negatives=20)
    assert float(out0["loss"].item()) == pytest.approx(ref, rel=1e-5 if precision == "fp32" else 2e-2)
    full = oracle.retrieval_loss(qn, cn, temperature=0.5)
    assert ref < full                                                                # fewer negatives in the partition sum
    losses = [float(model.train_step(batch)["loss"].item()) for _ in range(6)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses      # descent along the tape's gradients
    with pytest.raises(NotImplementedError):
        tt.tasks.Retrieval(num_hard_negatives=5, remove_accidental_hits=True)(model.last_q, model.last_c, candidate_ids=np.arange(B))
    with pytest.raises(ValueError):
        tt.tasks.Retrieval(num_hard_negatives=0)
