"""f3 / f4: the evaluation loop's Recall@k / NDCG@k / MRR against a numpy restatement on the same
embeddings, and the three console entry points the reference declares (/root/reference/pyproject.toml:66-69)."""
import io
import json
import sys

import numpy as np
import pytest
import torch

import oracle
from two_tower_b200 import cli, evaluation, recipes, synth

pytestmark = pytest.mark.gpu


def _numpy_metrics(topk_ids, true_ids, ks):
    """Recall@k (= tfrs top_k_categorical_accuracy with one relevant item), NDCG@k = 1 / log2(rank + 2), MRR."""
    K = topk_ids.shape[1]
    ranks = np.full(len(true_ids), K, dtype=np.int64)
    for i, (row, t) in enumerate(zip(topk_ids, true_ids)):
        hit = np.nonzero(row == t)[0]
        if hit.size:
            ranks[i] = hit[0]
    out = {}
    for k in ks:
        h = ranks < min(k, K)
        out[f"recall@{k}"] = float(h.mean())
        out[f"ndcg@{k}"] = float(np.where(h, 1.0 / np.log2(ranks + 2.0), 0.0).mean())
    out["mrr"] = float(np.where(ranks < K, 1.0 / (ranks + 1.0), 0.0).mean())
    return out


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_evaluator_metrics_equal_the_oracle_on_the_same_embeddings(tt, precision):
    tt.set_precision(precision)
    cfg = synth.Config("eval", 99, 512, 64, 3000, 2500, (128, 64), 0.2)
    model = recipes.build_two_tower(cfg, lr=0.05)
    rng = synth.rng_for(12)
    batches = [synth.make_batch(cfg, s) for s in range(4)]
    model.test_step(batches[0])
    for b in batches:                                   # a few steps so that the positives rank above chance
        model.train_step(b)
    ks = (1, 5, 10, 20, 50, 100)
    ev = evaluation.RetrievalEvaluator(model.user_model, model.item_model, num_items=cfg.v_item, ks=ks)
    got = ev.evaluate(batches)
    # the same embeddings, pulled off the device, through the oracle's brute-force top-k (correctly rounded fp32 scores)
    items = model.item_model(torch.arange(cfg.v_item, dtype=torch.int64, device="cuda")).numpy()
    rows, truth = [], []
    for b in batches:
        q = model.user_model(b[recipes.USER_KEY]).numpy()
        _s, ids = oracle.brute_force_topk(q, items, 100, score_dtype=np.float32)
        rows.append(ids); truth.append(b[recipes.ITEM_KEY])
    ref = _numpy_metrics(np.concatenate(rows), np.concatenate(truth), ks)
    assert got["n_queries"] == 4 * cfg.batch
    for name, v in ref.items():
        assert got[name] == pytest.approx(v, abs=1e-12), name
    for k in ks:
        assert got[f"factorized_top_k/top_{k}_categorical_accuracy"] == got[f"recall@{k}"]
    assert got["recall@100"] > 2 * 100 / cfg.v_item         # above chance (100 / 2500) after four steps on these very pairs


def test_train_evaluate_serve_entry_points(tt, tmp_path, capsys, monkeypatch):
    # layer initialisers are seeded by (config.seed, process-wide layer counter): pin both so that the recall bound
    # below does not depend on which tests built layers before this one
    tt.set_seed(11)
    tt.layers._layer_counter[0] = 0
    cfg_yaml = tmp_path / "data_config.yaml"
    cfg_yaml.write_text(
        "model:\n  embedding_dim: 64\n  user_tower_dims: [128, 64]\n  item_tower_dims: [128, 64]\n  dropout_rate: 0.1\n"
        "  l2_regularization: 1.0e-6\n  training:\n    batch_size: 256\n    learning_rate: 0.05\n    epochs: 3\n    patience: 5\n"
        "    validation_freq: 1\n  retrieval:\n    candidate_sampling: \"in_batch\"\n    temperature: 0.1\n    top_k_eval: [1, 5, 10, 20, 50, 100]\n")
    ckpt = tmp_path / "model.npz"
    assert cli.main(["train", "--data", "synthetic:cfg1", "--config", str(cfg_yaml), "--out", str(ckpt)]) == 0
    rec = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert rec["epochs_run"] == 3 and len(rec["loss"]) == 3 and rec["loss"][-1] < rec["loss"][0]
    assert ckpt.exists() and rec["validation"]["recall@100"] > 0.2
    assert cli.main(["evaluate", "--data", "synthetic:cfg1", "--checkpoint", str(ckpt), "--config", str(cfg_yaml)]) == 0
    ev = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert ev["recall@100"] == pytest.approx(rec["validation"]["recall@100"], abs=0.05) and 0 < ev["mrr"] <= 1
    monkeypatch.setattr(sys, "stdin", io.StringIO("0 1 2 3 999\n"))
    assert cli.main(["serve", "--checkpoint", str(ckpt), "--config", str(cfg_yaml), "--k", "7"]) == 0
    out = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert out["user_ids"] == [0, 1, 2, 3, 999] and np.array(out["item_ids"]).shape == (5, 7)
    s = np.array(out["scores"])
    assert (np.diff(s, axis=1) <= 0).all()                  # sorted by score, descending
    assert cli.main([]) == 2
