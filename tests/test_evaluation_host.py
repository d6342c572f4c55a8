"""Host logic of the evaluation loop: rank bookkeeping, metric definitions, early stopping.  CPU only."""
import math

import pytest
import torch


def test_ranks_and_metrics(tt):
    from two_tower_b200 import evaluation as ev
    topk = torch.tensor([[5, 9, 2, 7], [1, 3, 8, 4], [6, 0, 0, 6], [2, 2, 2, 2]])
    true = torch.tensor([2, 9, 6, 2])
    ranks = ev.ranks_of_true_ids(topk, true)
    assert ranks.tolist() == [2, 4, 0, 0]                      # absent -> K; duplicates -> first position
    m = ev.metrics_from_ranks(ranks, 4, (1, 3, 100))
    assert m["recall@1"] == pytest.approx(2 / 4) and m["recall@3"] == pytest.approx(3 / 4) and m["recall@100"] == pytest.approx(3 / 4)
    assert m["ndcg@1"] == pytest.approx((1 + 1) / 4)
    assert m["ndcg@3"] == pytest.approx((1 / math.log2(4) + 1 + 1) / 4)
    assert m["mrr"] == pytest.approx((1 / 3 + 0 + 1 + 1) / 4)
    assert m["factorized_top_k/top_3_categorical_accuracy"] == m["recall@3"]
    assert ev.TOP_K_EVAL == (1, 5, 10, 20, 50, 100)            # /root/reference/configs/data_config.yaml:71


def test_early_stopping_patience(tt):
    from two_tower_b200 import evaluation as ev
    es = ev.EarlyStopping(monitor="recall@10", patience=2)
    seq = [0.10, 0.20, 0.19, 0.20, 0.18]
    stops = [es.update(i, {"recall@10": v}) for i, v in enumerate(seq)]
    assert stops == [False, False, False, True, True] and es.best == 0.20 and es.best_epoch == 1
    lo = ev.EarlyStopping(monitor="loss", patience=1, mode="min")
    assert [lo.update(i, {"loss": v}) for i, v in enumerate([3.0, 2.0, 2.5])] == [False, False, True]
    with pytest.raises(KeyError):
        es.update(9, {"other": 1.0})
    with pytest.raises(ValueError):
        ev.EarlyStopping(mode="up")


def test_dense_bucket_layout(tt):
    """Offsets of the flat dense-gradient bucket (every variable padded to a multiple of 4 floats: 16-byte vector loads)."""
    offs, total = tt.ops.bucket_layout([(128, 256), (1, 256), (256, 128), (1, 130), (3,)])
    assert offs == [0, 32768, 33024, 65792, 65924] and total == 65928
    assert all(o % 4 == 0 for o in offs) and total % 4 == 0


def test_model_spec_defaults_are_the_reference_config(tmp_path):
    """cli.load_model_spec: the `model:` block of the reference's configs/data_config.yaml:54-71 (defaults when no file
    is given) and overrides from a file."""
    from two_tower_b200 import cli
    s = cli.load_model_spec(None)
    assert (s.embedding_dim, s.user_tower_dims, s.item_tower_dims) == (128, (512, 256, 128), (512, 256, 128))
    assert (s.batch_size, s.learning_rate, s.epochs, s.patience, s.validation_freq) == (1024, 0.001, 50, 5, 1)
    assert s.temperature == 0.1 and s.top_k_eval == (1, 5, 10, 20, 50, 100) and s.l2 == 1e-6
    p = tmp_path / "c.yaml"
    p.write_text("model:\n  embedding_dim: 64\n  training:\n    batch_size: 256\n  retrieval:\n    temperature: 0.5\n")
    s = cli.load_model_spec(str(p))
    assert (s.embedding_dim, s.batch_size, s.temperature, s.epochs) == (64, 256, 0.5, 50)
    p.write_text("model:\n  retrieval:\n    candidate_sampling: uniform\n")
    import pytest
    with pytest.raises(NotImplementedError):
        cli.load_model_spec(str(p))
