"""The two-tower models of the BASELINE.json configurations, built from the TFRS-shaped layers.

The reference declares its model only as the `model:` block of /root/reference/configs/data_config.yaml:54-71
(`embedding_dim`, `user_tower_dims` / `item_tower_dims`, `temperature`, in-batch sampling) over the id columns its
data layer emits (/root/reference/src/data/preprocessor.py:481-489: `user_id_encoded`, `item_id_encoded`,
`category_encoded`); `src/models` itself is empty.  `build_two_tower(cfg)` is that model for a `synth.Config`:

    user tower : Embedding(v_user, d)                                         -> Dense stack
    item tower : Embedding(v_item, d) [+ mean-pooled EmbeddingBag per multi-hot feature, summed] -> Dense stack
    task       : tasks.Retrieval(temperature)

bench.py, the entry points (cli.py), smoke() and the parity tests all build their models here, so that a config name
means the same model everywhere.
"""
from __future__ import annotations

from typing import Optional

from . import layers, optimizers, tasks
from .models import Model

USER_KEY = "user_id_encoded"
ITEM_KEY = "item_id_encoded"


def item_feature_keys(cfg):
    return [ITEM_KEY, *cfg.bags.keys()]


def make_optimizer(name: str, lr: float):
    name = name.lower()
    if name == "adagrad":
        return optimizers.Adagrad(learning_rate=lr)
    if name in ("lazy_adam", "lazyadam"):
        return optimizers.LazyAdam(learning_rate=lr)
    raise ValueError(f"optimizer must be 'adagrad' or 'lazy_adam' (row-wise table updates), got {name!r}")


class TwoTower(Model):
    """tfrs.models.Model subclass: compute_loss(features) = task(user_model(user ids), item_model(item features))."""

    def __init__(self, cfg, pooling: str = "mean"):
        super().__init__()
        self.cfg = cfg

        def mlp():
            return [layers.Dense(u, "relu" if j < len(cfg.mlp) - 1 else None) for j, u in enumerate(cfg.mlp)]

        self.user_model = layers.Sequential([layers.Embedding(cfg.v_user, cfg.dim, name="user_embedding"), *mlp()])
        if cfg.bags:
            feats = {ITEM_KEY: layers.Embedding(cfg.v_item, cfg.dim, name="item_embedding")}
            for name, (vocab, _lmin, _lmax) in cfg.bags.items():
                feats[name] = layers.EmbeddingBag(vocab, cfg.dim, combiner=pooling, name=f"{name}_embedding")
            first = layers.FeatureSum(feats, name="item_features")
        else:
            first = layers.Embedding(cfg.v_item, cfg.dim, name="item_embedding")
        self.item_model = layers.Sequential([first, *mlp()])
        self.task = tasks.Retrieval(temperature=cfg.temperature)

    def item_inputs(self, features):
        if self.cfg.bags:
            return {k: features[k] for k in item_feature_keys(self.cfg)}
        return features[ITEM_KEY]

    def compute_loss(self, features, training: bool = False):
        return self.task(self.user_model(features[USER_KEY]), self.item_model(self.item_inputs(features)))


def build_two_tower(cfg, lr: float = 0.001, optimizer: str = "adagrad", pooling: str = "mean") -> TwoTower:
    model = TwoTower(cfg, pooling)
    model.compile(optimizer=make_optimizer(optimizer, lr))
    return model


def algorithmic_gather_bytes(cfg, avg_bag_len: Optional[dict] = None) -> int:
    """SURVEY.md 8(d) K1 bytes of ONE step, both towers: nnz * d * 4 (fp32 table rows read) + B * d * 2 per tower (bf16
    tower input written) + nnz * 8 (ids) + (B + 1) * 8 per bag feature (offsets)."""
    b, d = cfg.batch, cfg.dim
    nnz_user = b
    nnz_item = b
    extra = 0
    for name, (_vocab, lmin, lmax) in cfg.bags.items():
        mean_len = (avg_bag_len or {}).get(name, (lmin + lmax) / 2.0)
        nnz_item += int(round(b * mean_len))
        extra += (b + 1) * 8
    return (nnz_user + nnz_item) * (d * 4 + 8) + 2 * b * d * 2 + extra
