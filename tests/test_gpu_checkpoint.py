"""Raw-tensor checkpoint: a restored model continues bit-identically (tables, Dense weights, bf16
shadows, Adagrad accumulators, iteration count)."""
import numpy as np
import pytest
import torch

from two_tower_b200 import synth

pytestmark = pytest.mark.gpu


def _model(tt, opt):
    class TwoTower(tt.models.Model):
        def __init__(s):
            super().__init__()
            s.user_model = tt.Sequential([tt.layers.Embedding(900, 128), tt.layers.Dense(256, "relu"), tt.layers.Dense(128)])
            s.item_model = tt.Sequential([tt.layers.Embedding(700, 128), tt.layers.Dense(256, "relu"), tt.layers.Dense(128)])
            s.task = tt.tasks.Retrieval(temperature=0.3)

        def compute_loss(s, f, training=False):
            return s.task(s.user_model(f["user_id_encoded"]), s.item_model(f["item_id_encoded"]))

    m = TwoTower()
    m.compile(optimizer=opt)
    return m


@pytest.mark.parametrize("opt_name", ["adagrad", "lazy_adam"])
def test_restored_model_continues_identically(tt, tmp_path, opt_name):
    tt.set_precision("bf16")
    mk = (lambda: tt.optimizers.Adagrad(0.05)) if opt_name == "adagrad" else (lambda: tt.optimizers.LazyAdam(0.01))
    rng = synth.rng_for(77)
    batches = [{"user_id_encoded": synth.draw_ids(rng, 384, 900, 1.2), "item_id_encoded": synth.draw_ids(rng, 384, 700, 1.2)} for _ in range(4)]
    a = _model(tt, mk())
    a.test_step(batches[0])
    for b in batches[:2]:
        a.train_step(b)
    path = tmp_path / "ckpt.npz"
    a.save_weights(path)
    b_model = _model(tt, mk())
    b_model.test_step(batches[0])                     # build (different random init), then restore
    b_model.load_weights(path)
    assert b_model.optimizer.iterations == a.optimizer.iterations == 2
    for va, vb in zip(a.trainable_variables, b_model.trainable_variables):
        assert torch.equal(va.value, vb.value), va.name                    # restored bit for bit
        for k, s in va.slots.items():
            if isinstance(s, torch.Tensor) and not k.startswith("_"):
                assert torch.equal(s, vb.slots[k]), (va.name, k)
    for bt in batches[2:]:
        la, lb = float(a.train_step(bt)["loss"].item()), float(b_model.train_step(bt)["loss"].item())
        assert la == pytest.approx(lb, rel=1e-6)
    # the two runs continue together (duplicate ids are summed with fp32 atomics: the order, hence the last bits, may differ)
    for va, vb in zip(a.trainable_variables, b_model.trainable_variables):
        assert torch.allclose(va.value, vb.value, rtol=1e-4, atol=1e-6), va.name
    other = _model(tt, mk())
    with pytest.raises(ValueError):
        other.user_model = tt.Sequential([tt.layers.Embedding(901, 128), tt.layers.Dense(256, "relu"), tt.layers.Dense(128)])
        other.test_step({"user_id_encoded": batches[0]["user_id_encoded"], "item_id_encoded": batches[0]["item_id_encoded"]})
        other.load_weights(path)
