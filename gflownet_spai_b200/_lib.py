"""ctypes binding of libspai_b200.so (include/spai_b200.h).

The library is the product: there is no Python/CPU fallback. Loading fails
loudly when the shared object is missing (build it with
``python -c "import __graft_entry__ as g; g.build()"``), and every call raises ``SpaiError`` /
``ValueError`` on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspai_b200.so")

SPAI_OK, SPAI_ERR_INVALID, SPAI_ERR_CUDA, SPAI_ERR_UNSUPPORTED, SPAI_ERR_NOMEM = range(5)
MODE_COPY, MODE_LS = 0, 1
F32, F64 = 0, 1
MODE_LS_GRAM = 2
MODES = {"copy": MODE_COPY, "ls": MODE_LS, "ls_gram": MODE_LS_GRAM}

EXPORTS = [
    "spai_abi_version", "spai_last_error", "spai_device_count", "spai_ctx_create",
    "spai_ctx_destroy", "spai_ctx_info", "spai_ctx_set_workspace_limit",
    "spai_reward_batch_host", "spai_reward_batch_dev", "spai_kept_mask_dev",
    "spai_reward_from_taken_dev", "spai_row_index_sets", "spai_ls_solve_values_host",
    "spai_residual_pair_host", "spai_sample_step_dev", "spai_pack_taken_dev", "spai_reward_rows_dev", "spai_finalize_rewards_dev",
    "spai_ctx_enable_timing",
    "spai_ctx_last_timing", "spai_ctx_set_deletion_hint",
    "spai_reward_batch_host_len", "spai_reward_batch_dev_len",
    "spai_ingest_coo_to_csr_dev", "spai_ingest_spgemm_count_dev", "spai_ingest_spgemm_fill_dev",
    "spai_ingest_superset_dev", "spai_ingest_neumann_dev", "spai_ingest_csr_drop_zeros_dev",
    "spai_sample_taken_dev", "spai_sample_order_dev", "spai_sample_steps_dev", "spai_ctx_k3m_rows",
]


class SpaiError(RuntimeError):
    pass


class SpaiInfo(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("num_edges", C.c_int64), ("init_nnz", C.c_int64),
        ("num_actions", C.c_int64), ("a_nnz_stored", C.c_int64), ("a_nnz", C.c_int64),
        ("orig_flops", C.c_int64), ("orig_residual_f32", C.c_double),
        ("orig_residual_f64", C.c_double), ("contributions", C.c_int64),
        ("max_row_slots", C.c_int32), ("max_row_union", C.c_int32),
        ("has_duplicates", C.c_int32), ("device", C.c_int32),
        ("rows_missing_diag", C.c_int64), ("ls_class_rows", C.c_int64 * 8),
        ("device_bytes", C.c_int64),
    ]


class SpaiTiming(C.Structure):
    _fields_ = [
        ("ms_masks", C.c_float), ("ms_transpose", C.c_float), ("ms_reward", C.c_float),
        ("ms_finalize", C.c_float), ("ms_total", C.c_float), ("launches", C.c_int32),
        ("chunks", C.c_int32), ("algorithmic_bytes", C.c_double),
        ("compulsory_bytes", C.c_double), ("h2d_bytes", C.c_double),
    ]


_lib = None


def load():
    """Load the shared library once; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise SpaiError(
            f"{LIB_PATH} not found: the CUDA library is the only implementation of the reward "
            "path (no CPU fallback). Build it with __graft_entry__.build().")
    lib = C.CDLL(LIB_PATH)
    p64, pd, pv = C.POINTER(C.c_int64), C.POINTER(C.c_double), C.c_void_p
    i64, dbl, i32 = C.c_int64, C.c_double, C.c_int
    lib.spai_abi_version.restype = i32
    lib.spai_last_error.restype = C.c_char_p
    lib.spai_device_count.argtypes = [C.POINTER(i32)]
    lib.spai_ctx_create.argtypes = [i32, i64, i64, pv, pv, pv, i64, pv, pv, pv, C.POINTER(pv)]
    lib.spai_ctx_destroy.argtypes = [pv]
    lib.spai_ctx_destroy.restype = None
    lib.spai_ctx_info.argtypes = [pv, C.POINTER(SpaiInfo)]
    lib.spai_ctx_set_workspace_limit.argtypes = [pv, i64]
    lib.spai_reward_batch_host.argtypes = [pv, pv, i64, i64, i64, dbl, i32, i32, pv, pv, pv, pv]
    lib.spai_reward_batch_dev.argtypes = [pv, pv, i64, i64, i64, dbl, i32, i32, pv, pv, pv, pv]
    lib.spai_reward_batch_host_len.argtypes = [pv, pv, i32, pv, i64, i64, i64, dbl, i32, i32, pv, pv, pv, pv]
    lib.spai_reward_batch_dev_len.argtypes = [pv, pv, i32, pv, i64, i64, i64, dbl, i32, i32, pv, pv, pv, pv]
    lib.spai_ingest_coo_to_csr_dev.argtypes = [i32, i64, i64, pv, pv, pv, pv, pv, pv, p64, pv]
    lib.spai_ingest_spgemm_count_dev.argtypes = [i32, i64, pv, pv, pv, pv, pv, p64, pv]
    lib.spai_ingest_spgemm_fill_dev.argtypes = [i32, i64, pv, pv, pv, pv, pv, pv, pv, pv, pv, pv]
    lib.spai_ingest_csr_drop_zeros_dev.argtypes = [i32, i64, pv, pv, pv, pv, pv, pv, p64, pv]
    lib.spai_ingest_superset_dev.argtypes = [i32, i64, pv, pv, i32, i32, i32, pv, pv, pv, p64, pv]
    lib.spai_ingest_neumann_dev.argtypes = [i32, i64, pv, pv, pv, pv, pv, i32, pd, pv, pv]
    lib.spai_kept_mask_dev.argtypes = [pv, pv, i64, i64, i64, pv, pv]
    lib.spai_reward_from_taken_dev.argtypes = [pv, pv, i64, i64, dbl, i32, i32, pv, pv, pv, pv]
    lib.spai_row_index_sets.argtypes = [pv, i64, p64, pv, p64, pv]
    lib.spai_ls_solve_values_host.argtypes = [pv, pv, i64, i32, pv, pv]
    lib.spai_residual_pair_host.argtypes = [i32, i64, i64, pv, pv, pv, i64, pv, pv, pv, i32, pd, p64]
    lib.spai_sample_step_dev.argtypes = [pv, pv, i64, i64, pv, i64, pv, pv, i64, pv, pv, pv]
    lib.spai_reward_rows_dev.argtypes = [pv, pv, i64, i64, i64, i32, i32, i64, i64, pv, pv, pv]
    lib.spai_finalize_rewards_dev.argtypes = [pv, pv, pv, i64, dbl, i32, pv, pv, pv]
    lib.spai_pack_taken_dev.argtypes = [pv, pv, i64, i64, i64, pv, i64, pv, pv]
    lib.spai_sample_taken_dev.argtypes = [i32, pv, i64, i64, C.c_uint64, i64, pv, i64, pv, pv, i64, pv]
    lib.spai_sample_order_dev.argtypes = [i32, pv, i64, i64, C.c_uint64, i64, pv, pv, i32, i64, pv]
    lib.spai_sample_steps_dev.argtypes = [i32, pv, i64, i64, pv, i64, pv, pv, C.c_uint64, i64, i64, i64, pv, i32, pv, i64,
                                          pv, pv]
    lib.spai_ctx_k3m_rows.argtypes = [pv, p64, p64]
    lib.spai_ctx_set_deletion_hint.argtypes = [pv, i64]
    lib.spai_ctx_enable_timing.argtypes = [pv, i32]
    lib.spai_ctx_last_timing.argtypes = [pv, C.POINTER(SpaiTiming)]
    for name in EXPORTS:
        if name not in ("spai_last_error", "spai_ctx_destroy"):
            getattr(lib, name).restype = i32
    if lib.spai_abi_version() != 2:
        raise SpaiError(f"ABI mismatch: library reports {lib.spai_abi_version()}")
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status == SPAI_OK:
        return
    msg = load().spai_last_error().decode("utf-8", "replace")
    if status == SPAI_ERR_INVALID:
        raise ValueError(f"{what}: {msg}")
    raise SpaiError(f"{what}: status {status}: {msg}")
