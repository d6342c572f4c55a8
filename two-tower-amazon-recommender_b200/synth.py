"""Synthetic inputs of the BASELINE.json configurations.  Pure numpy host code;
every generator is seeded with numpy.random.Generator(PCG64(seed)).  The reference ships no
datasets reachable offline; id columns follow the schema the data layer emits
(/root/reference/src/data/preprocessor.py:481-489: dense int64 codes 0..V-1)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Tuple

import numpy as np


@dataclass
class Config:
    name: str
    seed: int
    batch: int
    dim: int
    v_user: int
    v_item: int
    mlp: Tuple[int, ...] = ()
    temperature: float | None = None
    zipf: float | None = None            # None -> uniform ids
    bags: Dict[str, Tuple[int, int, int]] = field(default_factory=dict)   # name -> (vocab, Lmin, Lmax)


CONFIGS = {
    # development_config sample-size 1000 (README.md:34-36): tiny, CPU-runnable
    "cfg1": Config("cfg1", 1234, 256, 64, 1000, 1000, (), 0.1, zipf=1.1),
    # single-B200 ID-only two-tower, 1M users x 500K items, d=128, B=8192, MLP 256-128
    "cfg2": Config("cfg2", 2345, 8192, 128, 1_000_000, 500_000, (256, 128), 0.1),
    # ID + multi-hot category/brand with mean pooling, 10M items, B=16384
    "cfg3": Config("cfg3", 3456, 16384, 128, 10_000_000, 10_000_000, (256, 128), 0.1,
                   bags={"category": (32768, 1, 8), "brand": (1_048_576, 1, 2)}),
    # row-sharded 100M-row tables across 8 GPUs, B_glob = 65536 (8192 per GPU), global-batch negatives
    "cfg4": Config("cfg4", 4567, 8192, 128, 100_000_000, 100_000_000, (256, 128), 0.1),
}


def rng_for(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def draw_ids(rng: np.random.Generator, n: int, vocab: int, zipf: float | None = None) -> np.ndarray:
    if zipf is None:
        return rng.integers(0, vocab, size=n, dtype=np.int64)
    z = rng.zipf(zipf, size=n).astype(np.int64) - 1
    return np.minimum(z, vocab - 1)


def draw_bags(rng: np.random.Generator, n: int, vocab: int, lmin: int, lmax: int, empty_frac: float = 0.0):
    lens = rng.integers(lmin, lmax + 1, size=n, dtype=np.int64)
    if empty_frac > 0:
        lens[rng.random(n) < empty_frac] = 0
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    values = rng.integers(0, vocab, size=int(offsets[-1]), dtype=np.int64)
    return values, offsets


def make_batch(cfg: Config, step: int = 0, batch: int | None = None, pad_bags: bool = False) -> Dict[str, object]:
    """One batch of the config: user_id_encoded / item_id_encoded (+ bag features as CSR (values, offsets)).
    pad_bags: `values` padded with -1 to the static capacity batch * Lmax (CUDA-graph replay needs fixed shapes; the
    kernels read offsets[batch] members and drop negative ids)."""
    rng = rng_for(cfg.seed * 1_000_003 + step)
    b = batch or cfg.batch
    out = {
        "user_id_encoded": draw_ids(rng, b, cfg.v_user, cfg.zipf),
        "item_id_encoded": draw_ids(rng, b, cfg.v_item, cfg.zipf),
    }
    for name, (vocab, lmin, lmax) in cfg.bags.items():
        values, offsets = draw_bags(rng, b, vocab, lmin, lmax)
        if pad_bags:
            padded = np.full(b * lmax, -1, dtype=np.int64)
            padded[:values.size] = values
            values = padded
        out[name] = (values, offsets)
    return out


def exact_matrix(rng: np.random.Generator, n: int, d: int, levels: int = 4) -> np.ndarray:
    """Entries in {-levels..levels}/8: every product and partial sum of a d<=4096 dot product
    is exactly representable in fp32 (and the entries in bf16), so scores are independent of
    accumulation order and ties are real -- used for bit-exact top-k id tests."""
    return (rng.integers(-levels, levels + 1, size=(n, d)).astype(np.float32)) / 8.0
