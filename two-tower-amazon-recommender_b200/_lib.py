"""ctypes binding of libtwotower.so (include/twotower.h).  No fallback: if the library is
missing or a call fails, this raises."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libtwotower.so"

TT_F32, TT_BF16 = 0, 1
TT_POOL_SUM, TT_POOL_MEAN = 0, 1
TT_MAX_FEATURES = 8


class TwoTowerError(RuntimeError):
    """A libtwotower call returned a negative status; carries tt_last_error()."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libtwotower error {code}: {message}")
        self.code = code
        self.message = message


class tt_feature(C.Structure):
    _fields_ = [("table", C.c_void_p), ("values", C.c_void_p), ("offsets", C.c_void_p),
                ("vocab", C.c_int64), ("mode", C.c_int32), ("shard_world", C.c_int32)]


TT_MAX_TOWERS, TT_MAX_DENSE_VARS, TT_MAX_SPARSE_VARS = 4, 16, 8


class tt_tower_mlp2(C.Structure):
    _fields_ = [("feats", tt_feature * TT_MAX_FEATURES), ("num_feats", C.c_int32), ("d_in", C.c_int32),
                ("d_hid", C.c_int32), ("d_out", C.c_int32), ("batch", C.c_int64),
                ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
                ("x", C.c_void_p), ("h", C.c_void_p), ("y", C.c_void_p),
                ("dy_parts", C.c_void_p), ("dy_splits", C.c_int32), ("reserved", C.c_int32),
                ("dx", C.c_void_p), ("dw1_parts", C.c_void_p), ("dw2_parts", C.c_void_p),
                ("db1_parts", C.c_void_p), ("db2_parts", C.c_void_p),
                ("prepare_workspace", C.c_void_p), ("prepare_workspace_bytes", C.c_int64)]


class tt_dense_var(C.Structure):
    _fields_ = [("w", C.c_void_p), ("slot0", C.c_void_p), ("slot1", C.c_void_p), ("grad_parts", C.c_void_p),
                ("n", C.c_int64), ("num_parts", C.c_int32), ("l2", C.c_float), ("shadow", C.c_void_p)]


class tt_sparse_var(C.Structure):
    _fields_ = [("table", C.c_void_p), ("slot0", C.c_void_p), ("slot1", C.c_void_p),
                ("vocab", C.c_int64), ("d", C.c_int64), ("values", C.c_void_p), ("offsets", C.c_void_p),
                ("num_rows", C.c_int64), ("nnz", C.c_int64), ("grad", C.c_void_p), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_int64), ("first_flag", C.c_void_p), ("mode", C.c_int32),
                ("shard", C.c_int32)]


_p, _i64, _i32, _f = C.c_void_p, C.c_int64, C.c_int32, C.c_float

# symbol -> (restype, argtypes); mirrors include/twotower.h one to one
SIGNATURES = {
    "tt_version": (C.c_int, []),
    "tt_last_error": (C.c_char_p, []),
    "tt_device_check": (C.c_int, []),
    "tt_profile_enable": (C.c_int, [_i32]),
    "tt_profile_collect": (_i64, [C.c_char_p, _i64]),
    "tt_tower_input_fwd": (C.c_int, [C.POINTER(tt_feature), _i32, _p, _p, _i64, _i64, _p, _p]),
    "tt_embedding_gather_f32": (C.c_int, [_p, _p, _p, _i64, _i64, _i64, _p]),
    "tt_embedding_gather_bf16": (C.c_int, [_p, _p, _p, _i64, _i64, _i64, _p]),
    "tt_embedding_bag_fwd": (C.c_int, [_p, _p, _p, _i32, _p, _i32, _i64, _i64, _i64, _p]),
    "tt_sparse_workspace_bytes": (_i64, [_i64, _i64]),
    "tt_sparse_workspace_init": (C.c_int, [_p, _i64, _i64, _i64, _p]),
    "tt_sparse_adagrad_update": (C.c_int, [_p, _p, _i64, _i64, _p, _p, _i32, _i64, _i64, _p, _f, _f, _p, _i64, _p, _p]),
    "tt_sparse_lazy_adam_update": (C.c_int, [_p, _p, _p, _i64, _i64, _p, _p, _i32, _i64, _i64, _p, _f, _f, _f, _f, _p, _i64, _p, _p]),
    "tt_dense_adagrad_update": (C.c_int, [_p, _p, _p, _i32, _i64, _i64, _f, _f, _f, _p, _p]),
    "tt_dense_adam_update": (C.c_int, [_p, _p, _p, _p, _i32, _i64, _i64, _f, _f, _f, _f, _f, _p, _p]),
    "tt_sum_squares": (C.c_int, [_p, _i64, _f, _p, _i32, _p]),
    "tt_dense_adagrad_update_multi": (C.c_int, [C.POINTER(tt_dense_var), _i32, _f, _f, _p]),
    "tt_dense_adam_update_multi": (C.c_int, [C.POINTER(tt_dense_var), _i32, _f, _f, _f, _f, _p]),
    "tt_sparse_adagrad_update_multi": (C.c_int, [C.POINTER(tt_sparse_var), _i32, _f, _f, _p]),
    "tt_sparse_lazy_adam_update_multi": (C.c_int, [C.POINTER(tt_sparse_var), _i32, _f, _f, _f, _f, _p]),
    "tt_optimizer_prepare_sparse": (C.c_int, [C.POINTER(tt_sparse_var), _i32, _p]),
    "tt_fold_parts_multi": (C.c_int, [C.POINTER(tt_dense_var), _i32, _p]),
    "tt_adagrad_step": (C.c_int, [C.POINTER(tt_dense_var), _i32, C.POINTER(tt_sparse_var), _i32, _f, _f, _p]),
    "tt_lazy_adam_step": (C.c_int, [C.POINTER(tt_dense_var), _i32, C.POINTER(tt_sparse_var), _i32, _f, _p, _f, _f, _f, _p]),
    "tt_adam_bias_correction": (C.c_int, [_p, _f, _f, _f, _p, _p]),
    "tt_tower_mlp2_supported": (_i32, [_i32, _i32, _i32]),
    "tt_tower_mlp2_fwd": (C.c_int, [C.POINTER(tt_tower_mlp2), _i32, _p, _p]),
    "tt_tower_mlp2_bwd": (C.c_int, [C.POINTER(tt_tower_mlp2), _i32, _p]),
    "tt_hard_negative_loss_fwd": (C.c_int, [_i32, _p, _p, _p, _i64, _i64, _i64, _i64, _f, _p, _p, _p, _p, _p, _p, _p]),
    "tt_hard_negative_loss_bwd": (C.c_int, [_i32, _p, _p, _p, _i64, _i64, _i64, _i64, _f, _f, _p, _p, _p, _p, _p, _p]),
    "tt_retrieval_fwd_dq_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "tt_retrieval_loss_fwd_dq": (C.c_int, [_p, _p, _i64, _i64, _i64, _f, _i64, _p, _p, _p, _p, _p, _p, _i64, _p, _p]),
    "tt_retrieval_loss_bwd_dc_fused": (C.c_int, [_p, _p, _i64, _i64, _i64, _f, _i64, _p, _p, _f, _p, _p]),
    "tt_retrieval_bwd_num_splits": (C.c_int, [_i32, _i64, _i64, _i64, C.POINTER(_i32), C.POINTER(_i32)]),
    "tt_retrieval_loss_bwd_parts": (C.c_int, [_i32, _p, _p, _i64, _i64, _i64, _f, _i64, _p, _p, _p, _p, _f, _p, _p, _p]),
    "tt_combine_parts_f32": (C.c_int, [_p, _i32, _i64, _i64, _p, _p, _p]),
    "tt_dense_fwd": (C.c_int, [_i32, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i32, _p]),
    "tt_dense_bwd": (C.c_int, [_i32, _p, _p, _p, _p, _p, _p, _i32, _p, _i64, _i64, _i64, _i32, _p]),
    "tt_dense_bwd_num_parts": (_i32, [_i32, _i64, _i64, _i64]),
    "tt_colsum_f32": (C.c_int, [_p, _p, _i64, _i64, _i32, _p]),
    "tt_sum_parts_f32": (C.c_int, [_p, _i32, _i64, _p, _p]),
    "tt_debug_gemm_bf16": (C.c_int, [_p, _i32, _p, _i32, _i64, _i64, _i64, _p, _p]),
    "tt_cast_f32_to_bf16": (C.c_int, [_p, _p, _i64, _p]),
    "tt_debug_trace_buffer": (C.c_int, [_p]),
    "tt_debug_tower_trace": (C.c_int, [_p]),
    "tt_debug_timeline": (C.c_int, [_p]),
    "tt_debug_topk_scan_mode": (C.c_int, [_i32]),
    "tt_debug_topk_scan_trace": (C.c_int, [_p]),
    "tt_retrieval_workspace_bytes": (_i64, [_i32, _i64, _i64, _i64]),
    "tt_retrieval_workspace_init": (C.c_int, [_i32, _p, _i64, _i64, _i64, _i64, _p]),
    "tt_retrieval_loss_fwd": (C.c_int, [_i32, _p, _p, _i64, _i64, _i64, _f, _i64, _p, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "tt_retrieval_loss_bwd": (C.c_int, [_i32, _p, _p, _i64, _i64, _i64, _f, _i64, _p, _p, _p, _p, _f,
                                         _p, _p, _p, _p, _p, _i64, _p]),
    "tt_topk_num_splits": (_i32, [_i32, _i64, _i64, _i64, _i32]),
    "tt_topk_num_launches": (_i32, [_i32, _i64, _i64, _i64, _i32]),
    "tt_topk_workspace_bytes": (_i64, [_i32, _i64, _i64, _i64, _i32]),
    "tt_topk_bruteforce": (C.c_int, [_i32, _p, _p, _i64, _i64, _i64, _i32, _i64, _p, _p, _p, _p, _p, _i64, _p]),
    "tt_topk_bruteforce_peer": (C.c_int, [_i32, _p, _p, _i64, _i64, _i64, _i32, _i64, _p, _i32, _i32, _i64, _i64, _i64, _p, _p, _i64, _p]),
    "tt_topk_merge": (C.c_int, [_p, _p, _i32, _i64, _i32, _i32, _i64, _p, _p, _p, _p]),
    "tt_topk_hits": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i32, C.POINTER(_i32), _i32, _p, _p, _p]),
    "tt_rowwise_dot": (C.c_int, [_i32, _p, _p, _p, _i64, _i64, _p]),
    "tt_partition_ids": (C.c_int, [_p, _i64, _i32, _i64, _p, _p, _p, _p, _p]),
    "tt_permute_rows": (C.c_int, [_p, _p, _p, _i64, _i64, _i32, _p]),
    "tt_peer_barrier": (C.c_int, [_p, _p, _i32, _i32, _i32, _p]),
    "tt_peer_push": (C.c_int, [_p, _i32, _i32, C.POINTER(_p), C.POINTER(_i64), C.POINTER(_i64), _p]),
    "tt_peer_sum_f32": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _p, _i32, _i32, _i32, _p]),
    "tt_peer_combine_scatter": (C.c_int, [_p, _p, _i32, _i64, _i64, _i64, _i64, _i32, _i32, _p]),
    "tt_peer_make_row_maps": (C.c_int, [C.POINTER(C.c_uint64), _i32, _i64, _i64, _p]),
    "tt_peer_retrieval_bwd_dc": (C.c_int, [_p, _p, _i64, _i64, _i64, _f, _i64, _p, _p, _f, _p, _i32, _i32, _p, _p]),
    "tt_peer_push_rows": (C.c_int, [_p, _i32, C.POINTER(_p), C.POINTER(_p), C.POINTER(_i64), _i64, _i64, _i32, _i32, _p]),
    "tt_peer_pull_rows": (C.c_int, [_p, _i32, C.POINTER(_p), C.POINTER(_i64), C.POINTER(_p), _i64, _i64, _p, _p,
                                    _i32, _i32, _i32, _p]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen libtwotower.so and bind every symbol of include/twotower.h.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} not found.  Build it with `python __graft_entry__.py build` (nvcc, sm_100a). "
            "There is no CPU or PyTorch fallback for the two-tower hot path.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().tt_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != 0:
        raise TwoTowerError(rc, last_error())
