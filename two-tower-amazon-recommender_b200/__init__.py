"""two_tower_b200 -- B200-native (sm_100a) two-tower training + retrieval hot path behind a
TensorFlow-Recommenders-shaped Python surface.  See DESIGN.md / INTEGRATION.md.

    import two_tower_b200 as tt
    user_model = tt.Sequential([tt.layers.Embedding(V, 128), tt.layers.Dense(256, "relu"), tt.layers.Dense(128)])
    task = tt.tasks.Retrieval(temperature=0.1)
    class TwoTower(tt.models.Model):
        def compute_loss(self, features, training=False):
            return self.task(self.user_model(features["user_id_encoded"]),
                             self.item_model(features["item_id_encoded"]))
"""
from . import _lib, core, data, evaluation, layers, metrics, models, ops, optimizers, recipes, tasks  # noqa: F401
from . import serving  # noqa: F401,E402
from ._lib import TwoTowerError  # noqa: F401
from .core import GradientTape, Tensor, Variable, config, set_precision, set_seed  # noqa: F401
from .layers import Dense, Embedding, EmbeddingBag, FeatureSum, Sequential  # noqa: F401

__version__ = "0.2.0"
