"""The C-ABI library loads, exports every symbol include/twotower.h declares, and validates
arguments without touching a GPU (CPU-only checks)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "twotower.h"
DEBUG_HEADER = ROOT / "include" / "twotower_debug.h"      # measurement / test hooks, not the product ABI


def declared_symbols(headers=(HEADER, DEBUG_HEADER)):
    out = set()
    for h in headers:
        text = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        out |= set(re.findall(r"\b(tt_[a-z0-9_]+)\s*\(", text))
    return sorted(out)


def test_product_header_has_no_debug_hooks():
    assert not [s for s in declared_symbols((HEADER,)) if s.startswith(("tt_debug_", "tt_profile_"))]


def test_header_declares_the_path():
    syms = declared_symbols()
    for needed in ("tt_tower_input_fwd", "tt_dense_fwd", "tt_dense_bwd", "tt_retrieval_loss_fwd",
                   "tt_retrieval_loss_bwd", "tt_sparse_adagrad_update", "tt_topk_bruteforce", "tt_topk_merge"):
        assert needed in syms


def test_library_exports_every_declared_symbol(tt):
    lib = tt._lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in twotower.h but not exported"
        assert name in tt._lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(tt._lib.SIGNATURES) == set(declared_symbols())
    assert lib.tt_version() == 120


def test_argument_validation_reports_through_last_error(tt):
    lib = tt._lib.load()
    rc = lib.tt_embedding_gather_f32(None, None, None, 4, 6, 10, None)      # d % 4 != 0 and null table
    assert rc == -1
    assert b"tt_tower_input_fwd" in lib.tt_last_error()
    rc = lib.tt_retrieval_loss_fwd(0, 16, 16, 8, 4, 8, 1.0, 0, None, None, None, 16, 16, 16, None, 0, None)
    assert rc == -1 and b"exceed" in lib.tt_last_error()                   # labels past the candidates
    rc = lib.tt_topk_bruteforce(0, 16, 16, 4, 8, 8, 9, 0, None, 16, 16, None, None, 0, None)
    assert rc == -1 and b"k=9" in lib.tt_last_error()                      # k > num candidates
    with pytest.raises(tt.TwoTowerError) as e:
        tt._lib.check(lib.tt_topk_merge(None, None, 1, 1, 1, 1, 0, None, None, None, None))
    assert e.value.code == -1


def test_ops_refuse_cpu_tensors(tt):
    import torch
    with pytest.raises(TypeError, match="CUDA tensors only"):
        tt.ops.embedding_gather(torch.zeros(4, 8), torch.zeros(2, dtype=torch.int64))


def test_no_oracle_import_in_product():
    pkg = ROOT / "two-tower-amazon-recommender_b200"
    for py in pkg.rglob("*.py"):
        src = py.read_text()
        assert "import oracle" not in src and "from oracle" not in src, py
