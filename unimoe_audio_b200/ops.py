"""Thin torch-facing wrappers over the C ABI (device memory + streams only; no compute in Python).

Each function mirrors one stage of ``UniMoEAudioSparseMoeBlock.forward`` (reference
utils/UniMoE_Audio_core.py:236-358); see include/dcmoe_b200.h for the reference lines each replaces.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import DcmoeConfig, DcmoePlanLayout, DcmoeSizes, DcmoeWorkspace

_TORCH_DT = {torch.float32: _lib.DCMOE_F32, torch.bfloat16: _lib.DCMOE_BF16}


@dataclass(frozen=True)
class LayerDims:
    hidden_size: int = 2048
    n_real: int = 8
    n_null: int = 1
    n_fix: int = 2
    dynamic_intermediate_size: int = 2752
    shared_intermediate_size: int = 1376
    top_p: float = 0.7
    jitter_eps: float = 0.01
    fixed_top_k: int = 0      # mlp_dynamic_top_k, used when top_p == 0 (core.py:256-257)

    @property
    def n_dyn(self) -> int:
        return self.n_real + self.n_null

    @property
    def n_experts(self) -> int:
        return self.n_dyn + self.n_fix

    def c_config(self, dtype: torch.dtype) -> DcmoeConfig:
        if dtype not in _TORCH_DT:
            raise TypeError(f"DCMoE supports float32 and bfloat16, got {dtype}")
        return DcmoeConfig(self.hidden_size, self.n_real, self.n_null, self.n_fix, self.dynamic_intermediate_size,
                           self.shared_intermediate_size, _TORCH_DT[dtype], int(self.fixed_top_k) if self.top_p == 0 else 0,
                           float(self.top_p), float(self.jitter_eps))


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class on_device:
    """Make ``device`` the current CUDA device for the launches inside the block (the C side launches on the current
    device, and the stream handed to it must belong to that device).  Costs nothing when it already is current --
    the reference's PyTorch ops guard the device the same way, so a layer that lives on cuda:1 works from a process
    whose current device is cuda:0."""

    __slots__ = ("idx", "prev")

    def __init__(self, device):
        device = torch.device(device)
        self.idx = device.index if device.index is not None else torch.cuda.current_device()
        self.prev = -1

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev >= 0:
            torch.cuda.set_device(self.prev)
        return False


def guarded(fn):
    """Run ``fn`` with the device of its workspace / first tensor argument current."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = None
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, Workspace):
                dev = a.device
                break
            if isinstance(a, torch.Tensor) and a.is_cuda and dev is None:
                dev = a.device
        if dev is None:
            return fn(*args, **kwargs)
        with on_device(dev):
            return fn(*args, **kwargs)
    return wrapper


def query_sizes(dims: LayerDims, dtype: torch.dtype, T: int, row_capacity: int = 0):
    lib = _lib.load()
    cfg = dims.c_config(dtype)
    sz, lay = DcmoeSizes(), DcmoePlanLayout()
    _lib.check(lib.dcmoe_query_sizes(cfg, T, row_capacity, sz, lay), "dcmoe_query_sizes")
    return sz, lay


class Workspace:
    """Device buffers of one forward: the plan (counts / prefix sums / tile table / aux), the packed
    rows, the FFN intermediates and the permutation maps.  Sized for the worst case (every token routed
    to every real expert) unless ``row_capacity`` is given."""

    def __init__(self, dims: LayerDims, dtype: torch.dtype, T: int, device, row_capacity: int = 0,
                 alloc_peer_visible: bool = True):
        self.dims, self.dtype, self.T, self.device = dims, dtype, T, torch.device(device)
        self.sizes, self.layout = query_sizes(dims, dtype, T, row_capacity)
        self.cfg = dims.c_config(dtype)          # cached ctypes struct (built once per workspace)
        self.row_capacity = int(self.sizes.row_capacity)
        self.t_pad = int(self.sizes.t_pad)
        dev = self.device
        self.plan = torch.zeros(int(self.sizes.plan_bytes), dtype=torch.uint8, device=dev)
        H, Id = dims.hidden_size, dims.dynamic_intermediate_size
        self.shapes = {"x_packed": (max(self.row_capacity - self.t_pad, 1), H), "y": (self.row_capacity, H),
                       "row_scale": (self.row_capacity, 2)}
        self.h = torch.empty((self.row_capacity, Id), dtype=dtype, device=dev)
        if alloc_peer_visible:   # expert parallelism allocates these three with dcmoe_ipc_alloc instead
            self.x_packed = torch.empty(self.shapes["x_packed"], dtype=dtype, device=dev)
            self.y = torch.empty(self.shapes["y"], dtype=dtype, device=dev)
            self.row_scale = torch.zeros(self.shapes["row_scale"], dtype=torch.float32, device=dev)
        self.slot_of = torch.empty((max(T, 1), dims.n_real), dtype=torch.int32, device=dev)
        self.row_token = torch.full((self.row_capacity,), -1, dtype=torch.int32, device=dev)
        self._c_ws = None
        # a row_capacity below the worst case can overflow (rows dropped, plan.overflow = 1): the flag is copied to pinned
        # host memory after every forward and examined, without blocking, before the next one
        worst = self.t_pad + dims.n_real * T + _lib.TILE_M * dims.n_real
        self.reduced = self.row_capacity < worst
        self._ovf_host = None
        self._ovf_event = None

    def c_workspace(self) -> DcmoeWorkspace:
        """The dcmoe_workspace struct of this workspace's buffers (built once; the buffers never move)."""
        if self._c_ws is None:
            self._c_ws = DcmoeWorkspace(self.plan.data_ptr(), self.x_packed.data_ptr(), self.slot_of.data_ptr(),
                                        self.row_token.data_ptr(), self.row_scale.data_ptr(), self.h.data_ptr(),
                                        self.y.data_ptr())
        return self._c_ws

    # typed views into the plan buffer (device tensors; reading them on the host synchronises)
    def _view(self, off: int, n: int, dt: torch.dtype) -> torch.Tensor:
        return self.plan[off: off + n * 4].view(dt)

    @property
    def counts(self) -> torch.Tensor:
        return self._view(self.layout.counts, self.dims.n_real, torch.int32)

    @property
    def seg_base(self) -> torch.Tensor:
        return self._view(self.layout.seg_base, self.dims.n_real + 1, torch.int32)

    @property
    def n_mtiles(self) -> torch.Tensor:
        return self._view(self.layout.n_mtiles, 1, torch.int32)

    @property
    def overflowed(self) -> bool:
        """True if the last plan did not fit `row_capacity` (rows were dropped).  Reading it synchronises; the
        default worst-case capacity can never overflow."""
        return bool(self._view(self.layout.overflow, 1, torch.int32).item())

    def note_overflow_async(self):
        """Enqueue the copy of the plan's overflow flag to pinned host memory (no synchronisation)."""
        if self._ovf_host is None:
            self._ovf_host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self._ovf_event = torch.cuda.Event()
        self._ovf_host.copy_(self._view(self.layout.overflow, 1, torch.int32), non_blocking=True)
        self._ovf_event.record(torch.cuda.current_stream(self.device))

    def raise_if_overflowed(self, block: bool = False):
        """Raise if a forward that used this workspace dropped rows because ``row_capacity`` was too small.  With
        ``block=False`` only flags whose copy has already completed are examined."""
        if self._ovf_event is None:
            return
        if block:
            self._ovf_event.synchronize()
        elif not self._ovf_event.query():
            return
        if int(self._ovf_host[0]) != 0:
            self._ovf_host[0] = 0
            raise _lib.DcmoeError(
                f"DCMoE workspace overflow: the routed rows of a forward over {self.T} tokens did not fit row_capacity="
                f"{self.row_capacity}; rows were dropped and that call's output is incomplete.  Raise row_capacity / "
                f"row_capacity_factor (worst case: every token to every routed expert)")

    @property
    def aux_loss(self) -> torch.Tensor:
        return self._view(self.layout.aux_loss, 1, torch.float32)

    @property
    def mtiles(self) -> torch.Tensor:
        n = int(self.sizes.max_mtiles)
        return self._view(self.layout.mtiles, n * 4, torch.int32).view(n, 4)


@guarded
def router(x: Optional[torch.Tensor], w_gate: Optional[torch.Tensor], ws: Workspace, logits_in: Optional[torch.Tensor] = None,
           attention_mask: Optional[torch.Tensor] = None, keep: Optional[torch.Tensor] = None, fp32_gate: bool = False,
           all_fp32: bool = False):
    """Top-P router.  Returns (full_router_logits, dynamic_top_k, expert_mask, global_weight).  ``keep`` [T, E] uint8:
    token_drop's capacity mask (``drop_select``); ``fp32_gate``: fp32 logits and routing arithmetic on a bf16 layer
    (training-mode forward, core.py:240-249); ``all_fp32``: x / w_gate are float32 tensors and every output is float32
    although the workspace belongs to a bf16 layer (the fp32 gate on a jittered float copy of the input)."""
    lib = _lib.load()
    dims, T, dt, dev = ws.dims, ws.T, ws.dtype, ws.device
    E = dims.n_experts
    if all_fp32:
        dt, fp32_gate = torch.float32, False
    ldt = torch.float32 if fp32_gate else dt
    logits = torch.empty((T, E), dtype=ldt, device=dev)
    top_k = torch.empty((T,), dtype=torch.int64, device=dev)
    mask = torch.empty((T, E), dtype=torch.int32, device=dev)
    gw = torch.empty((T, E), dtype=dt, device=dev)
    am = None
    if attention_mask is not None:
        am = attention_mask.reshape(-1).to(device=dev, dtype=torch.int32).contiguous()
        if am.numel() != T:
            raise ValueError("attention_mask must have one entry per token")
    if logits_in is not None:
        if logits_in.shape != (T, E) or logits_in.dtype != ldt or not logits_in.is_contiguous():
            raise ValueError("logits_in must be a contiguous [T, E] tensor of the layer dtype (float32 with the fp32 gate)")
    if keep is not None and (keep.shape != (T, E) or keep.dtype != torch.uint8 or not keep.is_contiguous()):
        raise ValueError("keep must be a contiguous uint8 [T, E] tensor")
    cfg = ws.cfg if dt == ws.dtype else dims.c_config(dt)
    if keep is None and not fp32_gate:
        _lib.check(lib.dcmoe_router(_ptr(x), _ptr(w_gate), _ptr(logits_in), _ptr(am), T, cfg, _ptr(logits), _ptr(top_k),
                                    _ptr(mask), _ptr(gw), _ptr(ws.plan), _stream()), "dcmoe_router")
    else:
        _lib.check(lib.dcmoe_router_ex(_ptr(x), _ptr(w_gate), _ptr(logits_in), _ptr(am), _ptr(keep),
                                       _lib.ROUTER_FP32_GATE if fp32_gate else 0, T, cfg, _ptr(logits), _ptr(top_k),
                                       _ptr(mask), _ptr(gw), _ptr(ws.plan), _stream()), "dcmoe_router_ex")
    return logits, top_k, mask, gw


def expert_capacity(dims: LayerDims, T: int, capacity_factor: float, min_capacity: int) -> int:
    """core.py:170-175 + the clamp of :306-308 (host arithmetic, done on the C side in float32 as the reference does)."""
    import ctypes

    lib = _lib.load()
    cap = ctypes.c_int64(0)
    _lib.check(lib.dcmoe_expert_capacity(T, dims.c_config(torch.float32), float(capacity_factor), int(min_capacity),
                                         ctypes.byref(cap)), "dcmoe_expert_capacity")
    return int(cap.value)


@guarded
def drop_select(logits: torch.Tensor, expert_mask: torch.Tensor, capacity: int, ws: Workspace) -> torch.Tensor:
    """Capacity mask of token_drop / "probs" (core.py:305-314): uint8 [T, E]."""
    lib = _lib.load()
    T, E = logits.shape
    keep = torch.empty((T, E), dtype=torch.uint8, device=logits.device)
    scratch = getattr(ws, "_drop_keys", None)
    if scratch is None or scratch.numel() < ws.dims.n_dyn * max(T, 1):
        scratch = ws._drop_keys = torch.empty(ws.dims.n_dyn * max(T, 1), dtype=torch.int64, device=logits.device)
    _lib.check(lib.dcmoe_drop_select(_ptr(logits), _TORCH_DT[logits.dtype], _ptr(expert_mask), T, ws.cfg, int(capacity),
                                     _ptr(scratch), _ptr(keep), _stream()), "dcmoe_drop_select")
    return keep


@guarded
def aux_weighted(logits: torch.Tensor, expert_mask: torch.Tensor, aux_balance_weight: Optional[torch.Tensor], ws: Workspace) -> torch.Tensor:
    """Load-balancing loss with aux_balance_weight (core.py:380-385; None: the plain means of :378-379 evaluated in the
    logits' dtype): float32 0-dim tensor."""
    lib = _lib.load()
    T = logits.shape[0]
    integer, w = False, None
    if aux_balance_weight is not None:
        if aux_balance_weight.numel() != T:
            # core.py:381-383 broadcasts a [B, S] weight over num_hidden_layers = rows / (B * S) stacked layers; a single
            # block call always has rows == B * S
            raise ValueError("aux_balance_weight must have one entry per token ([batch, seq])")
        integer = not aux_balance_weight.dtype.is_floating_point
        w = aux_balance_weight.reshape(-1).to(device=logits.device, dtype=torch.float32).contiguous()
    scratch = getattr(ws, "_aux_scratch", None)
    n = (max(T, 1) + _lib.ROUTER_BLOCK - 1) // _lib.ROUTER_BLOCK * 32
    if scratch is None or scratch.numel() < n:
        scratch = ws._aux_scratch = torch.empty(n, dtype=torch.float32, device=logits.device)
    # (no tokens: the reference's means over an empty dimension are 0 / 0)
    aux = torch.full((), float("nan"), dtype=torch.float32, device=logits.device) if T == 0 else \
        torch.empty((), dtype=torch.float32, device=logits.device)
    _lib.check(lib.dcmoe_aux_weighted(_ptr(logits), _TORCH_DT[logits.dtype], _ptr(expert_mask), _ptr(w), 1 if integer else 0,
                                      T, ws.cfg, _ptr(scratch), _ptr(aux), _stream()), "dcmoe_aux_weighted")
    return aux


@guarded
def front_small(x: torch.Tensor, w_gate: torch.Tensor, ws: Workspace, attention_mask: Optional[torch.Tensor] = None):
    """router + plan + permute in one launch (bf16, T <= 64).  Returns what ``router`` returns."""
    lib = _lib.load()
    dims, T, dt, dev = ws.dims, ws.T, ws.dtype, ws.device
    E = dims.n_experts
    logits = torch.empty((T, E), dtype=dt, device=dev)
    top_k = torch.empty((T,), dtype=torch.int64, device=dev)
    mask = torch.empty((T, E), dtype=torch.int32, device=dev)
    gw = torch.empty((T, E), dtype=dt, device=dev)
    am = None
    if attention_mask is not None:
        am = attention_mask.reshape(-1).to(device=dev, dtype=torch.int32).contiguous()
        if am.numel() != T:
            raise ValueError("attention_mask must have one entry per token")
    _lib.check(lib.dcmoe_front_small(_ptr(x), _ptr(w_gate), _ptr(am), T, ws.row_capacity, ws.cfg, _ptr(logits), _ptr(top_k),
                                     _ptr(mask), _ptr(gw), _ptr(ws.plan), _ptr(ws.x_packed), _ptr(ws.slot_of),
                                     _ptr(ws.row_token), _ptr(ws.row_scale), _stream()), "dcmoe_front_small")
    return logits, top_k, mask, gw


@guarded
def plan(ws: Workspace):
    lib = _lib.load()
    _lib.check(lib.dcmoe_plan(ws.T, ws.row_capacity, ws.cfg, _ptr(ws.plan), _stream()), "dcmoe_plan")


@guarded
def permute(x: torch.Tensor, expert_mask: torch.Tensor, global_weight: torch.Tensor, ws: Workspace):
    lib = _lib.load()
    _lib.check(lib.dcmoe_permute(_ptr(x), _ptr(expert_mask), _ptr(global_weight), ws.T, ws.row_capacity,
                                 ws.cfg, _ptr(ws.plan), _ptr(ws.x_packed), _ptr(ws.slot_of),
                                 _ptr(ws.row_token), _ptr(ws.row_scale), _stream()), "dcmoe_permute")


@guarded
def grouped_ffn(x: torch.Tensor, w13: torch.Tensor, w2: torch.Tensor, ws: Workspace, impl: int = 0, phase: int = 0):
    lib = _lib.load()
    _lib.check(lib.dcmoe_grouped_ffn(_ptr(x), _ptr(ws.x_packed), _ptr(w13), _ptr(w2), _ptr(ws.row_scale), ws.T,
                                     ws.row_capacity, ws.cfg, _ptr(ws.plan), _ptr(ws.h), _ptr(ws.y),
                                     impl, phase, _stream()), "dcmoe_grouped_ffn")


@guarded
def combine(ws: Workspace, out: torch.Tensor, residual: Optional[torch.Tensor] = None,
            aux_out: Optional[torch.Tensor] = None):
    """Combine (+ optional fused residual add).  With ``aux_out`` (a float32 scalar tensor) the same launch also copies
    the plan's auxiliary loss into it, so the caller gets a per-call aux tensor without a separate copy kernel."""
    lib = _lib.load()
    if aux_out is None:
        _lib.check(lib.dcmoe_combine(_ptr(ws.y), _ptr(ws.slot_of), ws.T, ws.cfg, _ptr(residual), _ptr(out), _stream()),
                   "dcmoe_combine")
    else:
        _lib.check(lib.dcmoe_combine_aux(_ptr(ws.y), _ptr(ws.slot_of), ws.T, ws.cfg, _ptr(residual), _ptr(out),
                                         ws.plan.data_ptr() + ws.layout.aux_loss, _ptr(aux_out), _stream()),
                   "dcmoe_combine_aux")


@guarded
def forward(x: torch.Tensor, w_gate: torch.Tensor, w13: torch.Tensor, w2: torch.Tensor, ws: Workspace, out: torch.Tensor,
            attention_mask: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, impl: int = 0):
    """The whole layer in ONE host call (``dcmoe_forward``).  Returns (logits, top_k, mask, gw, aux)."""
    import ctypes

    lib = _lib.load()
    dims, T, dt, dev = ws.dims, ws.T, ws.dtype, ws.device
    E = dims.n_experts
    logits = torch.empty((T, E), dtype=dt, device=dev)
    top_k = torch.empty((T,), dtype=torch.int64, device=dev)
    mask = torch.empty((T, E), dtype=torch.int32, device=dev)
    gw = torch.empty((T, E), dtype=dt, device=dev)
    aux = torch.empty((), dtype=torch.float32, device=dev)
    am = None
    if attention_mask is not None:
        am = attention_mask.reshape(-1).to(device=dev, dtype=torch.int32).contiguous()
        if am.numel() != T:
            raise ValueError("attention_mask must have one entry per token")
    _lib.check(lib.dcmoe_forward(_ptr(x), _ptr(w_gate), _ptr(am), _ptr(w13), _ptr(w2), T, ws.row_capacity, ws.cfg,
                                 ctypes.byref(ws.c_workspace()), _ptr(residual), _ptr(out), _ptr(logits), _ptr(top_k),
                                 _ptr(mask), _ptr(gw), _ptr(aux), impl, _stream()), "dcmoe_forward")
    return logits, top_k, mask, gw, aux


def stream_segments(n_mtiles: int, granules_per_tile: int, grid: int):
    """Host mirror of ``cta_segment`` (csrc/ffn_tcgen05_stream.cu): the (m-tile, first granule, granule count) every CTA
    of the weight-streaming GEMMs works on.  CTAs are dealt to the m-tiles as evenly as possible, and the CTAs of one
    m-tile cut its granules evenly.  Used by the CPU tests to pin the partition's invariants."""
    out = []
    n_m = min(n_mtiles, grid)
    for c in range(grid):
        if n_m <= 0:
            out.append((0, 0, 0))
            continue
        base, extra = divmod(grid, n_m)
        if c < extra * (base + 1):
            m, idx, n = c // (base + 1), c % (base + 1), base + 1
        else:
            c2 = c - extra * (base + 1)
            m, idx, n = extra + c2 // base, c2 % base, base
        g0 = granules_per_tile * idx // n
        out.append((m, g0, granules_per_tile * (idx + 1) // n - g0))
    return out


@guarded
def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float, dims: LayerDims, out: Optional[torch.Tensor] = None):
    """Qwen2RMSNorm of [T, H] rows (the decoder layer's post_attention_layernorm, model.py:240)."""
    lib = _lib.load()
    if not (x.is_cuda and weight.is_cuda and x.is_contiguous() and weight.is_contiguous() and x.dtype == weight.dtype):
        raise ValueError("rmsnorm needs contiguous CUDA tensors of one dtype")
    if x.shape[-1] != dims.hidden_size or weight.numel() != dims.hidden_size:
        raise ValueError("rmsnorm: hidden size mismatch")
    out = torch.empty_like(x) if out is None else out
    T = x.numel() // dims.hidden_size
    _lib.check(lib.dcmoe_rmsnorm(_ptr(x), _ptr(weight), float(eps), T, dims.c_config(x.dtype), _ptr(out), _stream()),
               "dcmoe_rmsnorm")
    return out


@guarded
def pack_expert(gate_proj: torch.Tensor, up_proj: torch.Tensor, down_proj: torch.Tensor, group: int, part: int,
                dims: LayerDims, w13: torch.Tensor, w2: torch.Tensor):
    lib = _lib.load()
    dt = w13.dtype
    for t in (gate_proj, up_proj, down_proj):
        if t.dtype != dt or not t.is_contiguous() or not t.is_cuda:
            raise ValueError("expert weights must be contiguous CUDA tensors of the packed dtype")
    _lib.check(lib.dcmoe_pack_expert(_ptr(gate_proj), _ptr(up_proj), _ptr(down_proj), group, part, dims.c_config(dt),
                                     _ptr(w13), _ptr(w2), _stream()), "dcmoe_pack_expert")
