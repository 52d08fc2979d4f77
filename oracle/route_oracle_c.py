"""ctypes binding for oracle/route_oracle.c (test infrastructure only; see that file's header)."""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from . import build_oracle

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = build_oracle.build()
        _LIB = ctypes.CDLL(path)
        _LIB.dcmoe_oracle_route.restype = ctypes.c_int
        _LIB.dcmoe_oracle_route.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_double, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
        ]
        _LIB.dcmoe_oracle_route_k.restype = ctypes.c_int
        _LIB.dcmoe_oracle_route_k.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
        ]
        _LIB.dcmoe_oracle_route_ex.restype = ctypes.c_int
        _LIB.dcmoe_oracle_route_ex.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
            ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p,
        ]
        _LIB.dcmoe_oracle_exp_sleef.restype = ctypes.c_float
        _LIB.dcmoe_oracle_exp_sleef.argtypes = [ctypes.c_float]
        _LIB.dcmoe_oracle_exp_cr.restype = ctypes.c_float
        _LIB.dcmoe_oracle_exp_cr.argtypes = [ctypes.c_float]
        _LIB.dcmoe_oracle_exp_fast.restype = ctypes.c_float
        _LIB.dcmoe_oracle_exp_fast.argtypes = [ctypes.c_float, ctypes.POINTER(ctypes.c_int)]
    return _LIB


def route(logits: torch.Tensor, attention_mask: torch.Tensor | None = None, n_dyn: int = 9, n_fix: int = 2,
          top_p: float = 0.7, eps: float = 0.01, fixed_top_k: int = 0, keep: torch.Tensor | None = None,
          aux_weight: torch.Tensor | None = None):
    """Route ``logits`` [T, n_dyn+n_fix] (fp32 or bf16, CPU).

    Returns (dynamic_top_k int64 [T], expert_mask int32 [T,E], global_weight D [T,E], aux_loss fp32 0-dim),
    the 3rd..6th entries of the reference block's return tuple (utils/UniMoE_Audio_core.py:358).
    ``top_p == 0`` selects the fixed top-k branch (core.py:256-257) with ``fixed_top_k`` experts per token; the
    reference then returns dynamic_top_k as int32.
    """
    if top_p == 0 and fixed_top_k < 1:
        raise ValueError("top_p == 0 needs fixed_top_k >= 1 (mlp_dynamic_top_k)")
    fixed = int(fixed_top_k) if top_p == 0 else 0
    assert logits.dim() == 2 and logits.shape[1] == n_dyn + n_fix
    dt = logits.dtype
    assert dt in (torch.float32, torch.bfloat16)
    T, E = logits.shape
    lg = np.ascontiguousarray(logits.detach().cpu().float().numpy())
    am = None
    if attention_mask is not None:
        am = np.ascontiguousarray(attention_mask.detach().cpu().reshape(-1).to(torch.int32).numpy())
        assert am.shape[0] == T
    top_k = np.empty((T,), dtype=np.int64)
    mask = np.empty((T, E), dtype=np.int32)
    gw = np.empty((T, E), dtype=np.float32)
    aux = np.zeros((1,), dtype=np.float32)
    kp = aw = None
    if keep is not None:          # token_drop capacity mask [T, E] (core.py:313-316)
        kp = np.ascontiguousarray(keep.detach().cpu().to(torch.uint8).numpy())
        assert kp.shape == (T, E)
    if aux_weight is not None:    # aux_balance_weight [T] (core.py:380-385)
        aw = np.ascontiguousarray(aux_weight.detach().cpu().reshape(-1).to(torch.float32).numpy())
        assert aw.shape[0] == T
    rc = lib().dcmoe_oracle_route_ex(
        lg.ctypes.data, am.ctypes.data if am is not None else None, kp.ctypes.data if kp is not None else None,
        aw.ctypes.data if aw is not None else None, T, n_dyn, n_fix,
        1 if dt == torch.bfloat16 else 0, float(top_p), float(eps), fixed,
        top_k.ctypes.data, mask.ctypes.data, gw.ctypes.data, aux.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"dcmoe_oracle_route failed: {rc}")
    tk = torch.from_numpy(top_k)
    if fixed:
        tk = tk.to(torch.int32)
    return (tk, torch.from_numpy(mask), torch.from_numpy(gw).to(dt),
            torch.tensor(float(aux[0]), dtype=torch.float32))
