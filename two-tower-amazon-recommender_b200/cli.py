"""Console entry points the reference declares but does not ship:
/root/reference/pyproject.toml:66-69 names `train-model = src.training.train:main`, `evaluate-model =
src.evaluation.evaluate:main` and `serve-model = src.serving.api:main`; none of those modules exists.  These are thin
mains over the hot path -- `evaluation.fit`, `evaluation.RetrievalEvaluator`, `layers.factorized_top_k.BruteForce` --
parameterised by the `model:` block of the reference's /root/reference/configs/data_config.yaml:54-71.

    python -m two_tower_b200.cli train    --data combined_interactions.parquet --config data_config.yaml --out ckpt.npz
    python -m two_tower_b200.cli evaluate --data val.parquet --checkpoint ckpt.npz --config data_config.yaml
    python -m two_tower_b200.cli serve    --checkpoint ckpt.npz --config data_config.yaml --k 10 < user_ids.txt

`--data synthetic:cfg1` (or cfg2 / cfg3) draws the BASELINE synthetic interactions instead of reading a frame.  No HTTP
layer (FastAPI is out of scope): `serve` answers user ids read from stdin, one JSON line per query batch.
"""
from __future__ import annotations

import argparse
import json
import sys
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import data, evaluation, layers, optimizers, synth, tasks
from .core import set_precision
from .models import Model

# /root/reference/configs/data_config.yaml:54-71
DEFAULT_MODEL_BLOCK = {
    "embedding_dim": 128, "user_tower_dims": [512, 256, 128], "item_tower_dims": [512, 256, 128],
    "dropout_rate": 0.1, "l2_regularization": 1e-6,
    "training": {"batch_size": 1024, "learning_rate": 0.001, "epochs": 50, "patience": 5, "validation_freq": 1},
    "retrieval": {"candidate_sampling": "in_batch", "temperature": 0.1, "top_k_eval": [1, 5, 10, 20, 50, 100]},
}


@dataclass
class ModelSpec:
    embedding_dim: int = 128
    user_tower_dims: Tuple[int, ...] = (512, 256, 128)
    item_tower_dims: Tuple[int, ...] = (512, 256, 128)
    l2: float = 1e-6
    batch_size: int = 1024
    learning_rate: float = 0.001
    epochs: int = 50
    patience: int = 5
    validation_freq: int = 1
    temperature: Optional[float] = 0.1
    top_k_eval: Tuple[int, ...] = (1, 5, 10, 20, 50, 100)
    extra: Dict[str, object] = field(default_factory=dict)


def load_model_spec(path: Optional[str]) -> ModelSpec:
    """The `model:` block of the reference's data_config.yaml (defaults = the file's values).  dropout_rate is read and
    ignored: dropout draws random masks, which no parity run can reproduce; L2 applies to the Dense kernels."""
    block = dict(DEFAULT_MODEL_BLOCK)
    if path:
        import yaml
        with open(path) as f:
            doc = yaml.safe_load(f) or {}
        m = doc.get("model", {}) or {}
        for k, v in m.items():
            if isinstance(v, dict):
                block[k] = {**block.get(k, {}), **v}
            else:
                block[k] = v
    if block["retrieval"].get("candidate_sampling", "in_batch") != "in_batch":
        raise NotImplementedError("only candidate_sampling: in_batch is on the hot path")
    tr, re_ = block["training"], block["retrieval"]
    return ModelSpec(int(block["embedding_dim"]), tuple(int(x) for x in block["user_tower_dims"]),
                     tuple(int(x) for x in block["item_tower_dims"]), float(block.get("l2_regularization") or 0.0),
                     int(tr["batch_size"]), float(tr["learning_rate"]), int(tr["epochs"]), int(tr["patience"]),
                     int(tr["validation_freq"]), None if re_.get("temperature") is None else float(re_["temperature"]),
                     tuple(int(k) for k in re_["top_k_eval"]), {"dropout_rate": block.get("dropout_rate")})


class ConfiguredTwoTower(Model):
    """The two-tower model the reference's config describes: Embedding(V, embedding_dim) -> Dense stack per tower (relu on
    all but the last layer), in-batch softmax retrieval with temperature."""

    def __init__(self, spec: ModelSpec, num_users: int, num_items: int):
        super().__init__()
        self.spec = spec

        def tower(vocab, dims, name):
            seq = [layers.Embedding(vocab, spec.embedding_dim, name=f"{name}_embedding")]
            for j, u in enumerate(dims):
                seq.append(layers.Dense(u, "relu" if j < len(dims) - 1 else None, kernel_regularizer=spec.l2 or None,
                                        name=f"{name}_dense_{j}"))
            return layers.Sequential(seq, name=f"{name}_tower")
        self.user_model = tower(num_users, spec.user_tower_dims, "user")
        self.item_model = tower(num_items, spec.item_tower_dims, "item")
        self.task = tasks.Retrieval(temperature=spec.temperature)

    def compute_loss(self, features, training: bool = False):
        return self.task(self.user_model(features["user_id_encoded"]), self.item_model(features["item_id_encoded"]))


def _open_data(arg: str, batch_size: int, seed: int = 0):
    """-> (InteractionBatches, num_users, num_items)"""
    if arg.startswith("synthetic:"):
        cfg = synth.CONFIGS[arg.split(":", 1)[1]]
        rng = synth.rng_for(cfg.seed)
        n = max(4 * batch_size, 1000)
        frame = {"user_id_encoded": synth.draw_ids(rng, n, cfg.v_user, cfg.zipf),
                 "item_id_encoded": synth.draw_ids(rng, n, cfg.v_item, cfg.zipf)}
        ds = data.InteractionBatches(frame, batch_size=batch_size, seed=seed)
        return ds, cfg.v_user, cfg.v_item
    ds = data.InteractionBatches(arg, batch_size=batch_size, seed=seed)
    return ds, ds.num_users, ds.num_items


def _build(spec: ModelSpec, num_users: int, num_items: int, example) -> ConfiguredTwoTower:
    model = ConfiguredTwoTower(spec, num_users, num_items)
    model.compile(optimizer=optimizers.Adagrad(learning_rate=spec.learning_rate))
    model.test_step({k: v.cuda() for k, v in example.items()})          # builds the Dense layers
    return model


def _checkpoint_sizes(path: str) -> Tuple[int, int]:
    z = np.load(path, allow_pickle=False)
    names = [str(n) for n in z["names"]]
    u = next(i for i, n in enumerate(names) if n.startswith("user_embedding"))
    it = next(i for i, n in enumerate(names) if n.startswith("item_embedding"))
    return int(z[f"v{u}"].shape[0]), int(z[f"v{it}"].shape[0])


def train_main(argv: Optional[Sequence[str]] = None) -> int:
    ap = argparse.ArgumentParser(prog="train-model", description="Train the two-tower retrieval model (B200 hot path).")
    ap.add_argument("--data", required=True, help="interaction frame (parquet) or synthetic:cfg1|cfg2|cfg3")
    ap.add_argument("--validation", default=None, help="validation frame (parquet); default: the training frame")
    ap.add_argument("--config", default=None, help="the reference's data_config.yaml (model: block)")
    ap.add_argument("--epochs", type=int, default=None)
    ap.add_argument("--batch-size", type=int, default=None)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--out", default=None, help="checkpoint path (.npz)")
    a = ap.parse_args(argv)
    set_precision(a.precision)
    spec = load_model_spec(a.config)
    if a.batch_size:
        spec.batch_size = a.batch_size
    ds, nu, ni = _open_data(a.data, spec.batch_size)
    model = _build(spec, nu, ni, ds.example())
    val = ds if a.validation is None else data.InteractionBatches(a.validation, batch_size=spec.batch_size, shuffle=False)
    ks = tuple(k for k in spec.top_k_eval if k <= ni)
    ev = evaluation.RetrievalEvaluator(model.user_model, model.item_model, num_items=ni, ks=ks)
    step = model.train_step if a.no_graph else model.make_graphed_train_step({k: v.cuda() for k, v in ds.example().items()})
    hist = evaluation.fit(model, ds, epochs=a.epochs or spec.epochs, validation_batches=val, evaluator=ev,
                          validation_freq=spec.validation_freq,
                          early_stopping=evaluation.EarlyStopping(f"recall@{min(10, max(ks))}", patience=spec.patience), step=step)
    if a.out:
        model.save_weights(a.out)
    print(json.dumps({"epochs_run": len(hist["loss"]), "loss": hist["loss"], "stopped_epoch": hist["stopped_epoch"],
                      "validation": hist["val"][-1] if hist["val"] else None, "checkpoint": a.out}))
    return 0


def evaluate_main(argv: Optional[Sequence[str]] = None) -> int:
    ap = argparse.ArgumentParser(prog="evaluate-model", description="Recall@k / NDCG@k / MRR against the whole item corpus.")
    ap.add_argument("--data", required=True)
    ap.add_argument("--checkpoint", required=True)
    ap.add_argument("--config", default=None)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    a = ap.parse_args(argv)
    set_precision(a.precision)
    spec = load_model_spec(a.config)
    nu, ni = _checkpoint_sizes(a.checkpoint)
    ds, _nu, _ni = _open_data(a.data, spec.batch_size)
    ds.shuffle = False
    model = _build(spec, nu, ni, ds.example())
    model.load_weights(a.checkpoint)
    ks = tuple(k for k in spec.top_k_eval if k <= ni)
    ev = evaluation.RetrievalEvaluator(model.user_model, model.item_model, num_items=ni, ks=ks)
    print(json.dumps(ev.evaluate(ds)))
    return 0


def serve_main(argv: Optional[Sequence[str]] = None) -> int:
    ap = argparse.ArgumentParser(prog="serve-model", description="Exact top-k retrieval for user ids read from stdin.")
    ap.add_argument("--checkpoint", required=True)
    ap.add_argument("--config", default=None)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--input", default="-", help="file of whitespace-separated user ids (default stdin)")
    a = ap.parse_args(argv)
    set_precision(a.precision)
    spec = load_model_spec(a.config)
    nu, ni = _checkpoint_sizes(a.checkpoint)
    example = {"user_id_encoded": torch.zeros(1, dtype=torch.int64), "item_id_encoded": torch.zeros(1, dtype=torch.int64)}
    model = _build(spec, nu, ni, example)
    model.load_weights(a.checkpoint)
    index = evaluation.RetrievalEvaluator(model.user_model, model.item_model, num_items=ni, ks=(min(a.k, ni),)).build_index()
    text = sys.stdin.read() if a.input == "-" else open(a.input).read()
    ids = np.array([int(t) for t in text.split()], dtype=np.int64)
    if ids.size and (ids.min() < 0 or ids.max() >= nu):
        raise ValueError(f"user ids must be in [0, {nu})")
    for lo in range(0, ids.size, 4096):
        chunk = ids[lo:lo + 4096]
        scores, items = index(model.user_model(torch.from_numpy(chunk).cuda()), k=min(a.k, ni))
        print(json.dumps({"user_ids": chunk.tolist(), "item_ids": items.cpu().tolist(), "scores": scores.cpu().tolist()}))
    return 0


def main(argv: Optional[Sequence[str]] = None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    cmds = {"train": train_main, "evaluate": evaluate_main, "serve": serve_main}
    if not argv or argv[0] not in cmds:
        sys.stderr.write("usage: python -m two_tower_b200.cli {train|evaluate|serve} ...\n")
        return 2
    return cmds[argv[0]](argv[1:])


if __name__ == "__main__":
    raise SystemExit(main())
