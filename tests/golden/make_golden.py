"""Freeze oracle outputs at the cfg1 shape (B=256, d=64, V=1000, Zipf ids, T=0.1) into
tests/golden/cfg1.npz.  The reference repository has no implementation or vectors for this
path (PARITY UNPINNED, see oracle/__init__.py), so these vectors pin the ORACLE against drift
and are cross-checked here against independent torch CPU routines (cross_entropy,
EmbeddingBag, autograd) before being written.

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402
from two_tower_b200 import synth  # noqa: E402


def main():
    cfg = synth.CONFIGS["cfg1"]
    rng = synth.rng_for(cfg.seed)
    U = oracle.keras_uniform(rng, (cfg.v_user, cfg.dim))
    I = oracle.keras_uniform(rng, (cfg.v_item, cfg.dim))
    batch = synth.make_batch(cfg, 0)
    uid, iid = batch["user_id_encoded"], batch["item_id_encoded"]
    q = oracle.embedding_lookup(U, uid).astype(np.float64)
    c = oracle.embedding_lookup(I, iid).astype(np.float64)
    w = rng.uniform(0.5, 1.5, size=cfg.batch)
    p = rng.uniform(1e-4, 0.2, size=cfg.batch)
    out = {}
    for tag, kw in {
        "plain": dict(temperature=None),
        "temp": dict(temperature=cfg.temperature),
        "full": dict(temperature=cfg.temperature, sample_weight=w, candidate_sampling_probability=p,
                     candidate_ids=iid, remove_accidental_hits=True),
    }.items():
        r = oracle.retrieval_loss_and_grads(q, c, **kw)
        # independent check: torch cross_entropy + autograd
        tq = torch.tensor(q, requires_grad=True)
        tc = torch.tensor(c, requires_grad=True)
        s = torch.tensor(oracle.retrieval_scores(q, c, kw.get("temperature"), kw.get("candidate_sampling_probability"),
                                                 kw.get("candidate_ids"), kw.get("remove_accidental_hits", False)))
        s_t = tq @ tc.T
        if kw.get("temperature"):
            s_t = s_t / kw["temperature"]
        s_t = s_t + (s - s_t).detach()            # add the constant transforms
        ce = torch.nn.functional.cross_entropy(s_t, torch.arange(cfg.batch), reduction="none")
        wt = torch.tensor(kw.get("sample_weight", np.ones(cfg.batch)))
        loss_t = (ce * wt).sum()
        loss_t.backward()
        assert abs(loss_t.item() - r["loss"]) <= 1e-9 * abs(r["loss"]), (tag, loss_t.item(), r["loss"])
        assert np.allclose(tq.grad.numpy(), r["dq"], rtol=1e-9, atol=1e-12)
        assert np.allclose(tc.grad.numpy(), r["dc"], rtol=1e-9, atol=1e-12)
        out[f"{tag}_loss"] = np.float64(r["loss"])
        out[f"{tag}_lse"] = r["lse"]
        out[f"{tag}_dq"] = r["dq"].astype(np.float32)
        out[f"{tag}_dc"] = r["dc"].astype(np.float32)
    # sparse Adagrad on the user table with the 'temp' gradient
    r = oracle.retrieval_loss_and_grads(q, c, temperature=cfg.temperature)
    acc0 = np.full_like(U, 0.1, dtype=np.float64)
    t1, a1, uniq = oracle.adagrad_sparse(U.astype(np.float64), acc0, uid, r["dq"], lr=0.1)
    out["adagrad_unique_ids"] = uniq
    out["adagrad_rows"] = t1[uniq].astype(np.float32)
    out["adagrad_acc_rows"] = a1[uniq].astype(np.float32)
    # brute-force top-k of the batch queries against the whole item table
    s, ids = oracle.brute_force_topk(q, I.astype(np.float64), 100)
    out["topk_ids"] = ids.astype(np.int32)
    out["topk_scores"] = s.astype(np.float32)
    # bags: mean pooling with an empty bag, checked against torch EmbeddingBag
    vals, offs = synth.draw_bags(rng, 64, cfg.v_item, 0, 6)
    pooled = oracle.embedding_bag(I, vals, offs, "mean")
    eb = torch.nn.functional.embedding_bag(torch.tensor(vals), torch.tensor(I, dtype=torch.float64),
                                           torch.tensor(offs[:-1]), mode="mean")
    assert np.allclose(eb.numpy(), pooled, rtol=1e-12, atol=1e-15)
    out["bag_values"], out["bag_offsets"], out["bag_mean"] = vals, offs, pooled.astype(np.float32)
    out["uid"], out["iid"], out["w"], out["p"] = uid, iid, w, p
    np.savez_compressed(ROOT / "tests" / "golden" / "cfg1.npz", **out)
    print("wrote cfg1.npz:", {k: getattr(v, "shape", ()) for k, v in out.items()})


if __name__ == "__main__":
    main()
