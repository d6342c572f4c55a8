"""SURVEY.md 8 f1 on the GPU: a frame in the reference's on-disk schema drives the graphed training step from pinned
host batches (one H2D copy per step) and from device-resident id columns."""
import numpy as np
import pandas as pd
import pytest
import torch

pytestmark = pytest.mark.gpu


def _frame(n=4096, vu=300, vi=200, seed=9):
    rng = np.random.Generator(np.random.PCG64(seed))
    users = rng.integers(0, vu, n)
    items = (users * 7 + rng.integers(0, 3, n)) % vi            # learnable: the item follows from the user
    return pd.DataFrame({"user_idx": users.astype(np.int64), "item_idx": items.astype(np.int64),
                         "rating": np.ones(n, np.float32), "category": ["Books"] * n})


def test_training_from_a_parquet_in_the_reference_schema(tt, tmp_path):
    from two_tower_b200 import data
    tt.set_precision("bf16")
    tt.set_seed(3)
    path = tmp_path / "combined_interactions.parquet"
    _frame().to_parquet(path, index=False)
    ds = data.InteractionBatches(path, batch_size=512, seed=1)

    class TwoTower(tt.models.Model):
        def __init__(s):
            super().__init__()
            s.user_model = tt.Sequential([tt.layers.Embedding(ds.num_users, 128), tt.layers.Dense(256, "relu"), tt.layers.Dense(128)])
            s.item_model = tt.Sequential([tt.layers.Embedding(ds.num_items, 128), tt.layers.Dense(256, "relu"), tt.layers.Dense(128)])
            s.task = tt.tasks.Retrieval(temperature=0.2)

        def compute_loss(s, f, training=False):
            return s.task(s.user_model(f["user_id_encoded"]), s.item_model(f["item_id_encoded"]))

    model = TwoTower()
    model.compile(optimizer=tt.optimizers.Adagrad(0.05))
    example = {k: v.cuda() for k, v in ds.example().items()}
    model.test_step(example)
    step = model.make_graphed_train_step(example, warmup=2)
    losses = []
    for _ in range(3):
        epoch = [float(step(batch)["loss"].item()) for batch in ds]      # pinned host batches -> one H2D copy per step
        assert len(epoch) == len(ds) == 8 and all(np.isfinite(epoch))
        losses.append(np.mean(epoch))
    assert losses[-1] < losses[0]

    dev = ds.to_device()
    seen = []
    for batch in dev:
        assert all(t.is_cuda and t.dtype == torch.int64 and t.shape == (512,) for t in batch.values())
        seen.append(torch.stack([batch["user_idx"], batch["item_idx"]], 1))
        out = step(batch)                                                 # device batches feed the same graph
    assert np.isfinite(float(out["loss"].item()))
    seen = torch.cat(seen).cpu().numpy()
    ref = _frame()[["user_idx", "item_idx"]].to_numpy()
    # a permutation of the rows: every (user, item) pair drawn exists in the frame, no row twice
    key = lambda a: a[:, 0].astype(np.int64) * 100000 + a[:, 1]
    assert np.array_equal(np.sort(key(seen)), np.sort(key(ref)))
