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
    # The two runs continue together, each step from the SAME state (the second one after another save / restore).
    # Duplicate ids are summed with fp32 atomics, so the last bits of an updated row depend on the arrival order; carried
    # into a further step, a last-bit difference of a master weight can flip the rounding of its bf16 shadow or of an
    # activation (0.4 % of that element) and, with an Adagrad / Adam update as large as the weights themselves, grow to
    # several per cent of a table row -- between two runs of the SAME model just as well.  One step from identical state
    # is what "the restored model continues like the original" can be held to.
    for i, bt in enumerate(batches[2:]):
        if i > 0:
            a.save_weights(path)
            b_model.load_weights(path)
        la, lb = float(a.train_step(bt)["loss"].item()), float(b_model.train_step(bt)["loss"].item())
        assert la == pytest.approx(lb, rel=1e-6)       # bit-identical weights in, the same loss out
        for va, vb in zip(a.trainable_variables, b_model.trainable_variables):
            assert torch.allclose(va.value, vb.value, rtol=1e-5, atol=1e-7), (i, va.name)
        assert b_model.optimizer.iterations == a.optimizer.iterations == 3 + i
    other = _model(tt, mk())
    with pytest.raises(ValueError):
        other.user_model = tt.Sequential([tt.layers.Embedding(901, 128), tt.layers.Dense(256, "relu"), tt.layers.Dense(128)])
        other.test_step({"user_id_encoded": batches[0]["user_id_encoded"], "item_id_encoded": batches[0]["item_id_encoded"]})
        other.load_weights(path)


def test_graphed_lazy_adam_matches_eager_and_counts_iterations(tt):
    """Adam's bias-corrected step size changes every iteration: a captured graph must not freeze it.  The step count and
    alpha_t live on the device (tt_adam_bias_correction in the graph), optimizer.iterations follows the replays."""
    tt.set_precision("bf16")
    rng = synth.rng_for(91)
    mkb = lambda: {"user_id_encoded": torch.as_tensor(synth.draw_ids(rng, 384, 900)).cuda(),
                   "item_id_encoded": torch.as_tensor(synth.draw_ids(rng, 384, 700)).cuda()}
    b0 = mkb()
    a, b = _model(tt, tt.optimizers.LazyAdam(0.01)), _model(tt, tt.optimizers.LazyAdam(0.01))
    a.test_step(b0); b.test_step(b0)
    for va, vb in zip(a.trainable_variables, b.trainable_variables):
        vb.assign(va.numpy())
    graphed = b.make_graphed_train_step(b0, warmup=2)
    assert b.optimizer.iterations == 2
    for _ in range(2):
        a.train_step(b0)                                # mirror the two warm-up steps
    for s in range(6):
        bt = mkb()
        la, lb = float(a.train_step(bt)["loss"].item()), float(graphed(bt)["loss"].item())
        assert lb == pytest.approx(la, rel=2e-4 if s < 2 else 2e-3), s
    assert a.optimizer.iterations == b.optimizer.iterations == 8
    assert int(b.optimizer._dev_state[0].item()) == 8
    # alpha_8 on the device == the host formula at t = 8 (it would be alpha_3 if frozen at capture)
    # (betas are fp32 on the device, as in Keras: 1 - beta2^t carries their rounding -> 1e-5 relative)
    assert float(b.optimizer._dev_state[1].item()) == pytest.approx(b.optimizer._alpha(), rel=5e-5)
    frozen = tt.optimizers.LazyAdam(0.01); frozen.iterations = 3
    assert abs(float(b.optimizer._dev_state[1].item()) - frozen._alpha()) > 1e-4
    # six steps of two runs: atomics-order noise amplified by bf16 rounding flips and Adam's normalised update (see
    # above) can reach 1e-3 on single elements; a frozen alpha moves EVERY touched element by ~1e-2 over these steps
    for va, vb in zip(a.trainable_variables, b.trainable_variables):
        diff = (va.value - vb.value).abs()
        assert float(diff.max()) < 5e-3 and float(diff.mean()) < 1e-4, (va.name, float(diff.max()), float(diff.mean()))


def test_checkpoint_restore_is_seen_by_a_captured_graph(tt, tmp_path):
    """load_weights copies IN PLACE (values, bf16 shadows, optimizer slots): a graph captured before the restore
    trains the restored state, not stale buffers."""
    tt.set_precision("bf16")
    rng = synth.rng_for(92)
    mkb = lambda: {"user_id_encoded": torch.as_tensor(synth.draw_ids(rng, 384, 900)).cuda(),
                   "item_id_encoded": torch.as_tensor(synth.draw_ids(rng, 384, 700)).cuda()}
    b0, b1 = mkb(), mkb()
    m = _model(tt, tt.optimizers.Adagrad(0.05))
    m.test_step(b0)
    m.train_step(b0)
    path = tmp_path / "c.npz"
    m.save_weights(path)
    ptrs = [(v.value.data_ptr(), None if v.shadow is None else v.shadow.data_ptr(), v.slots["accumulator"].data_ptr())
            for v in m.trainable_variables]
    graphed = m.make_graphed_train_step(b0, warmup=2)
    for _ in range(3):
        graphed(b1)                                     # move away from the checkpoint
    m.load_weights(path)
    assert ptrs == [(v.value.data_ptr(), None if v.shadow is None else v.shadow.data_ptr(), v.slots["accumulator"].data_ptr())
                    for v in m.trainable_variables]
    ref = _model(tt, tt.optimizers.Adagrad(0.05))
    ref.test_step(b0)
    ref.load_weights(path)
    l_graph = float(graphed(b1)["loss"].item())
    l_ref = float(ref.train_step(b1)["loss"].item())
    assert l_graph == pytest.approx(l_ref, rel=1e-6)
    for va, vb in zip(m.trainable_variables, ref.trainable_variables):
        assert torch.allclose(va.value, vb.value, rtol=1e-4, atol=1e-6), va.name
