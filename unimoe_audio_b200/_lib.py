"""ctypes binding of libdcmoe_b200.so (the C ABI of include/dcmoe_b200.h).

There is no CPU fallback: if the shared library cannot be built/loaded, importing the ops fails
loudly, and every compute entry point returns DCMOE_ERR_CUDA when no CUDA device is present.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_double, c_float, c_int, c_int32, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdcmoe_b200.so")

DCMOE_F32, DCMOE_BF16 = 0, 1
ABI_VERSION = 3
ROUTER_FP32_GATE = 1
ROUTER_BLOCK, TILE_M = 16, 128


class DcmoeConfig(Structure):
    _fields_ = [
        ("hidden_size", c_int32), ("n_real", c_int32), ("n_null", c_int32), ("n_fix", c_int32),
        ("dynamic_intermediate_size", c_int32), ("shared_intermediate_size", c_int32),
        ("dtype", c_int32), ("fixed_top_k", c_int32), ("top_p", c_double), ("jitter_eps", c_double),
    ]


class DcmoeWorkspace(Structure):
    _fields_ = [("plan", c_void_p), ("x_packed", c_void_p), ("slot_of", c_void_p), ("row_token", c_void_p),
                ("row_scale", c_void_p), ("h", c_void_p), ("y", c_void_p)]


class DcmoeSizes(Structure):
    _fields_ = [("n_blocks", c_int64), ("t_pad", c_int64), ("max_mtiles", c_int64), ("row_capacity", c_int64),
                ("plan_bytes", c_int64)]


class DcmoePlanLayout(Structure):
    _fields_ = [(n, c_int64) for n in ("block_counts", "block_probs", "block_offsets", "counts", "seg_base",
                                       "n_mtiles", "aux_loss", "mtiles", "overflow", "total", "small_tokens")]


class DcmoeError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); must list every symbol include/dcmoe_b200.h declares
SIGNATURES = {
    "dcmoe_last_error": (ctypes.c_char_p, []),
    "dcmoe_abi_version": (c_int, []),
    "dcmoe_query_sizes": (c_int, [POINTER(DcmoeConfig), c_int64, c_int64, POINTER(DcmoeSizes), POINTER(DcmoePlanLayout)]),
    "dcmoe_router": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, POINTER(DcmoeConfig), c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p, c_void_p]),
    "dcmoe_router_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, POINTER(DcmoeConfig), c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dcmoe_test_exp": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "dcmoe_expert_capacity": (c_int, [c_int64, POINTER(DcmoeConfig), c_double, c_int64, POINTER(c_int64)]),
    "dcmoe_drop_select": (c_int, [c_void_p, c_int, c_void_p, c_int64, POINTER(DcmoeConfig), c_int64, c_void_p, c_void_p, c_void_p]),
    "dcmoe_aux_weighted": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, POINTER(DcmoeConfig), c_void_p, c_void_p,
                                   c_void_p]),
    "dcmoe_front_small": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, POINTER(DcmoeConfig), c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dcmoe_plan": (c_int, [c_int64, c_int64, POINTER(DcmoeConfig), c_void_p, c_void_p]),
    "dcmoe_permute": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, POINTER(DcmoeConfig), c_void_p, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_void_p]),
    "dcmoe_grouped_ffn": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                  POINTER(DcmoeConfig), c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "dcmoe_combine": (c_int, [c_void_p, c_void_p, c_int64, POINTER(DcmoeConfig), c_void_p, c_void_p, c_void_p]),
    "dcmoe_combine_aux": (c_int, [c_void_p, c_void_p, c_int64, POINTER(DcmoeConfig), c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p]),
    "dcmoe_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, POINTER(DcmoeConfig),
                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "dcmoe_rmsnorm": (c_int, [c_void_p, c_void_p, c_double, c_int64, POINTER(DcmoeConfig), c_void_p, c_void_p]),
    "dcmoe_pack_expert": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, POINTER(DcmoeConfig), c_void_p, c_void_p,
                                  c_void_p]),
    "dcmoe_ipc_alloc": (c_int, [c_int64, POINTER(c_void_p)]),
    "dcmoe_ipc_free": (c_int, [c_void_p]),
    "dcmoe_ipc_export": (c_int, [c_void_p, c_void_p]),
    "dcmoe_ipc_import": (c_int, [c_void_p, POINTER(c_void_p)]),
    "dcmoe_ipc_close": (c_int, [c_void_p]),
    "dcmoe_ep_plan": (c_int, [c_void_p, c_int, c_int, c_int64, c_int64, POINTER(DcmoeConfig), c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p]),
    "dcmoe_ep_dispatch": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, POINTER(DcmoeConfig), c_void_p, c_void_p,
                                  c_int, c_int, POINTER(c_void_p), POINTER(c_void_p), c_void_p, c_int, c_void_p]),
    "dcmoe_ep_combine": (c_int, [c_void_p, POINTER(c_void_p), c_void_p, c_int64, POINTER(DcmoeConfig), c_int, c_int,
                                 c_void_p, c_void_p, c_int, c_void_p]),
    "dcmoe_ep_barrier": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int32, c_void_p, c_int64, POINTER(c_void_p), c_void_p]),
    "dcmoe_ep_fetch_weights": (c_int, [POINTER(c_void_p), POINTER(c_void_p), c_int, c_int, POINTER(DcmoeConfig), c_void_p,
                                       c_void_p, c_void_p]),
}


def load(build_if_missing: bool = True):
    """Load (building first if the .so is missing or stale) and type the C ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        from . import build as _build
        try:
            _build.build()
        except Exception as exc:  # noqa: BLE001 - report and fail loudly below if no .so exists
            if not os.path.exists(LIB_PATH):
                raise DcmoeError(f"libdcmoe_b200.so is missing and could not be built: {exc}") from exc
    if not os.path.exists(LIB_PATH):
        raise DcmoeError(f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError = ABI mismatch, fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.dcmoe_abi_version() != ABI_VERSION:
        raise DcmoeError("libdcmoe_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().dcmoe_last_error().decode("utf-8", "replace")
        raise DcmoeError(f"{what} failed ({rc}): {msg}")
