"""Input-format bridge: frames shaped like the reference's data layer output ->
id batches.  CPU only (no pinning, no device)."""
import pickle

import numpy as np
import pandas as pd
import pytest
import torch


@pytest.fixture()
def frame():
    rng = np.random.Generator(np.random.PCG64(3))
    n = 1003
    users = rng.integers(0, 97, n); items = rng.integers(0, 211, n)
    return pd.DataFrame({
        "user_id": [f"U{u:04d}" for u in users], "parent_asin": [f"B{i:05d}" for i in items],
        "rating": rng.integers(1, 6, n).astype(np.float32), "timestamp": rng.integers(10**9, 2 * 10**9, n),
        "category": rng.choice(["Books", "Toys", "Music"], n),
        "user_idx": users.astype(np.int64), "item_idx": items.astype(np.int32),          # prepare_training_data.py:209-210
    })


def test_parquet_of_prepare_training_data_schema(tt, frame, tmp_path):
    from two_tower_b200 import data
    path = tmp_path / "combined_interactions.parquet"
    frame.to_parquet(path, compression="snappy", index=False)                         # as prepare_training_data.py:216
    ds = data.InteractionBatches(path, batch_size=128, seed=5, pin=False)
    assert (ds.num_rows, ds.num_users, ds.num_items) == (1003, int(frame.user_idx.max()) + 1, int(frame.item_idx.max()) + 1)
    assert len(ds) == 1003 // 128
    seen = []
    for batch in ds:
        assert set(batch) == {"user_idx", "item_idx", "user_id_encoded", "item_id_encoded"}
        assert all(t.dtype == torch.int64 and t.shape == (128,) for t in batch.values())
        assert torch.equal(batch["user_idx"], batch["user_id_encoded"]) and torch.equal(batch["item_idx"], batch["item_id_encoded"])
        seen.append(np.stack([batch["user_idx"].numpy().copy(), batch["item_idx"].numpy().copy()], 1))
    seen = np.concatenate(seen)
    order = ds.order(0)[:len(seen)]                                                    # the epoch's permutation, reproducible
    assert np.array_equal(seen[:, 0], frame.user_idx.to_numpy()[order]) and np.array_equal(seen[:, 1], frame.item_idx.to_numpy()[order])
    assert len(np.unique(order)) == len(order)                                          # every interaction at most once per epoch
    second = np.concatenate([b["user_idx"].numpy().copy() for b in ds])
    assert not np.array_equal(second, seen[:, 0])                                       # a new permutation per epoch
    again = data.InteractionBatches(path, batch_size=128, seed=5, pin=False)
    assert np.array_equal(np.concatenate([b["user_idx"].numpy().copy() for b in again]), seen[:, 0])   # same seed, same epoch 0


def test_preprocessor_schema_with_categories_and_remainder(tt, frame):
    from two_tower_b200 import data
    df = frame.rename(columns={"user_idx": "user_id_encoded", "item_idx": "item_id_encoded"})
    df["category_encoded"] = pd.factorize(df["category"])[0]                            # preprocessor.py:485-489
    ds = data.InteractionBatches(df, batch_size=250, shuffle=False, drop_remainder=False, pin=False)
    assert ds.num_categories == 3 and len(ds) == 5
    sizes = [b["category_encoded"].shape[0] for b in ds]
    assert sizes == [250, 250, 250, 250, 3]
    ex = ds.example()
    assert ex["user_id_encoded"].shape == (250,) and np.array_equal(ex["item_id_encoded"].numpy(), df.item_id_encoded.to_numpy()[:250])


def test_rejects_frames_without_encoded_ids_and_reads_mappings(tt, frame, tmp_path):
    from two_tower_b200 import data
    with pytest.raises(ValueError, match="no id columns"):
        data.InteractionBatches(frame[["user_id", "parent_asin", "rating"]], batch_size=8, pin=False)
    with pytest.raises(TypeError, match="integer ids"):
        data.InteractionBatches({"user_idx": np.array([0.5, 1.0]), "item_idx": np.array([1, 2])}, batch_size=1, pin=False)
    m = {"user_to_idx": {"U1": 0}, "item_to_idx": {"B1": 0}, "idx_to_user": {0: "U1"}, "idx_to_item": {0: "B1"}}
    p = tmp_path / "mappings.pkl"
    p.write_bytes(pickle.dumps(m))                                                      # prepare_training_data.py:222-232
    assert data.load_mappings(p)["item_to_idx"] == {"B1": 0}
    (tmp_path / "other.pkl").write_bytes(pickle.dumps({"x": 1}))
    with pytest.raises(ValueError, match="mappings.pkl"):
        data.load_mappings(tmp_path / "other.pkl")
