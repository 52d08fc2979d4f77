"""Expert-parallel DCMoE (reference: AudioMOELayer.forward with an ep_group, core.py:446-493; group wiring
core.py:505-520).

One process per GPU.  Rank r owns routed experts [r*n_loc, (r+1)*n_loc) (core.py:505), the gate and the shared experts
are replicated, every rank routes its own tokens.  Three paths, picked per call from the local token count T (every
rank of the group must call in lockstep with token counts on the same side of the two thresholds):

  decode    world * T <= 64.  Default policy "replicate": every rank keeps a RESIDENT copy of the remote experts' packs
            (fetched once, 270 MB x (world-1)/world per layer) and runs its tokens like a single GPU -- no exchange per
            call (72 us per layer call, the single-GPU latency).  Policy "exchange" (DCMOE_EP_DECODE=exchange; the same T
            on every rank): the TOKENS are replicated (pushed into every rank's buffer over NVLink), routed identically
            everywhere, each rank streams only ITS experts' weights, and the combine gathers each token's routed rows
            from their owners' y -- 88 us at 2 GPUs, 98 us at 8 (two cross-GPU barriers per call): at this size the
            layer is bound by per-kernel fixed costs, not by the weight bytes a rank streams.
  dispatch  the reference's exchange, un-padded: the permute kernel stores each selected row straight into the owner's
            packed buffer over NVLink (``ep_dispatch``), the owners run the grouped FFN on the rows they received, the
            combine kernel gathers the routed rows back with peer loads.
  gather    T >= gather_min_tokens: every token row that travels costs 2 x 4 KB (there and back) while the remote
            experts' weights are a fixed (world-1)/world x 270 MB per layer, so above ~9k tokens per rank it is the
            WEIGHTS that should travel: each rank pulls the remote experts' packs into a staging pack with copy-engine
            peer copies (no SM involved, double buffered so the fetch of call i+1 runs under the GEMMs of call i) and
            runs the single-GPU forward on its own tokens.  No dispatch, no combine exchange, no collective at all,
            and no load imbalance when the routing is skewed.  The experts stay sharded in HBM.

All cross-GPU synchronisation of the decode and dispatch paths is done through peer memory (``dcmoe_ep_barrier``:
release/acquire flags, optionally carrying the counts / token rows as a payload); NCCL is only used for the one-off
exchange of cudaIpc handles (``DCMOE_EP_FLAGS=0`` falls back to NCCL collectives).  Buffers touched by peers are
allocated with ``dcmoe_ipc_alloc`` ONCE per (group, device, dtype, layer dims) -- sized for ``max_tokens`` tokens per
rank -- and shared by all layers.  ``LocalRanks`` runs the same kernels for R virtual ranks inside one process on one
GPU (peer pointers are then ordinary local pointers); it is what the single-GPU tests use.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import replace
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib, ops
from .dcmoe import DCMoE
from .ops import LayerDims, Workspace

EP_META_INTS = 32
FLAG_SLOTS, MAX_RANKS = 8, 8
SLOT_COUNTS, SLOT_DISPATCH, SLOT_FFN, SLOT_DECODE_X, SLOT_DECODE_Y = 0, 1, 2, 3, 4


def ep_layout(all_counts: Sequence[Sequence[int]], rank: int, n_real: int):
    """Host mirror of ``ep_plan_kernel`` (csrc/ep.cu): from the all-gathered [world][n_real+1] table
    (per-rank rows per global expert, then the rank's token count) return
    (dest_base[e], dest_tpad[e], local_seg_base[l], local_totals[l]) for `rank`."""
    world = len(all_counts)
    n_loc = n_real // world
    pad = lambda v: (v + 127) // 128 * 128  # noqa: E731
    total = [sum(all_counts[r][e] for r in range(world)) for e in range(n_real)]
    before = [sum(all_counts[r][e] for r in range(rank)) for e in range(n_real)]
    dest_base, dest_tpad = [], []
    for e in range(n_real):
        owner = e // n_loc
        tp = pad(all_counts[owner][n_real])
        base = tp + sum(pad(total[q]) for q in range(owner * n_loc, e))
        dest_base.append(base + before[e])
        dest_tpad.append(tp)
    seg, row = [], pad(all_counts[rank][n_real])
    for l in range(n_loc):
        seg.append(row)
        row += pad(total[rank * n_loc + l])
    seg.append(row)
    return dest_base, dest_tpad, seg, [total[rank * n_loc + l] for l in range(n_loc)]


def choose_path(T: int, world: int, dtype, mode: str = "auto", gather_min_tokens: int = 8192, decode_ok: bool = True) -> str:
    """Which expert-parallel path a call with T local tokens takes (host logic, also used by the CPU tests)."""
    if T <= 0:
        return "dispatch"
    if decode_ok and dtype == torch.bfloat16 and T * world <= 64:
        return "decode"
    if mode == "gather" or (mode == "auto" and T >= gather_min_tokens):
        return "gather"
    return "dispatch"


class _RawCuda:
    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _wrap(ptr: int, nbytes: int, dtype, shape, device) -> torch.Tensor:
    return torch.as_tensor(_RawCuda(ptr, nbytes), device=device).view(dtype).view(shape)


def _ptr_array(values: Sequence[int]):
    return (ctypes.c_void_p * len(values))(*values)


class IpcBuffer:
    """Device memory from ``dcmoe_ipc_alloc`` (plain cudaMalloc: exportable with cudaIpcGetMemHandle, unlike a slice of
    the caching allocator's blocks).  ``ipc=False`` (virtual ranks, one process) uses the caching allocator."""

    def __init__(self, nbytes: int, device, ipc: bool, zero: bool = False):
        self.nbytes, self.device, self.ipc = int(nbytes), torch.device(device), ipc
        self._torch = None
        if ipc:
            p = ctypes.c_void_p()
            with ops.on_device(self.device):
                _lib.check(_lib.load().dcmoe_ipc_alloc(self.nbytes, ctypes.byref(p)), "dcmoe_ipc_alloc")
            self.ptr = p.value
        else:
            self._torch = torch.empty(self.nbytes, dtype=torch.uint8, device=self.device)
            self.ptr = self._torch.data_ptr()
        if zero:
            self.view(torch.uint8, (self.nbytes,)).zero_()

    def view(self, dtype, shape) -> torch.Tensor:
        return _wrap(self.ptr, self.nbytes, dtype, shape, self.device)

    def export(self) -> bytes:
        buf = (ctypes.c_uint8 * 64)()
        _lib.check(_lib.load().dcmoe_ipc_export(self.ptr, buf), "dcmoe_ipc_export")
        return bytes(buf)

    def free(self):
        if self.ipc and self.ptr:
            try:
                _lib.load().dcmoe_ipc_free(self.ptr)
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass
        self.ptr = 0
        self._torch = None


class EpWorkspace(Workspace):
    """Workspace whose peer-visible buffers (x_packed, y, row_scale) come from ``IpcBuffer``; it also owns the fp32
    partial sums of the overlapped combine, so that everything sized by T is rebuilt together."""

    def __init__(self, dims: LayerDims, dtype, T: int, device, row_capacity: int, ipc: bool, shared: Optional[dict] = None):
        super().__init__(dims, dtype, T, device, row_capacity, alloc_peer_visible=False)
        self.ep_meta = torch.zeros(EP_META_INTS, dtype=torch.int32, device=self.device)
        self.ipc = ipc
        es = torch.empty((), dtype=dtype).element_size()
        self._bufs: Dict[str, IpcBuffer] = {}
        for name, dt, esz in (("x_packed", dtype, es), ("y", dtype, es), ("row_scale", torch.float32, 4)):
            shape = self.shapes[name]
            if shared is not None and name in shared:       # views of buffers a larger workspace owns
                buf = shared[name]
                assert shape[0] * shape[1] * esz <= buf.nbytes
            else:
                buf = IpcBuffer(shape[0] * shape[1] * esz, self.device, ipc)
                self._bufs[name] = buf
            setattr(self, name, _wrap(buf.ptr, shape[0] * shape[1] * esz, dt, shape, self.device))
        self.row_scale.zero_()
        self.partial = torch.empty((max(T, 1), dims.hidden_size), dtype=torch.float32, device=self.device)

    def buffers(self) -> Dict[str, IpcBuffer]:
        return self._bufs

    def free(self):
        for b in self._bufs.values():
            b.free()
        self._bufs = {}

    def __del__(self):  # pragma: no cover - best effort
        self.free()


class EpContext:
    """What the layers of one model share per (expert-parallel group, device, dtype, layer dims): the peer-visible
    workspace of the dispatch path, the decode path's buffers, the flag array of the peer-memory barriers, the staging
    packs of the weight-gather path and the cudaIpc mappings of all of them.  Built collectively (every rank of the group
    at the same call); nothing in it is re-exchanged afterwards."""

    _registry: Dict[tuple, "EpContext"] = {}

    @classmethod
    def get(cls, group, rank: int, world: int, dims: LayerDims, dtype, device) -> "EpContext":
        key = (id(group) if group is not None else None, rank, world, dims, dtype, torch.device(device))
        ctx = cls._registry.get(key) if group is not None else None
        if ctx is None:
            ctx = cls(group, rank, world, dims, dtype, device)
            if group is not None:
                cls._registry[key] = ctx
        return ctx

    def __init__(self, group, rank: int, world: int, dims: LayerDims, dtype, device):
        self.group, self.rank, self.world = group, rank, world
        self.dims, self.dtype, self.device = dims, dtype, torch.device(device)
        self.n_loc = dims.n_real // world
        self.real = group is not None                  # separate processes (cudaIpc) vs virtual ranks in one process
        self.use_flags = os.environ.get("DCMOE_EP_FLAGS", "1") != "0"
        self._imported: List[int] = []
        # ---- dispatch path ----
        self.max_tokens = 0
        self.ws_full: Optional[EpWorkspace] = None     # owns the peer-visible buffers, sized for max_tokens per rank
        self.ws_by_T: Dict[int, EpWorkspace] = {}
        self.peer = None                               # (x_packed, row_scale, y) pointer arrays over the ranks
        self.counts_table: Optional[IpcBuffer] = None  # [world][n_real + 1] int32, filled by the peers (barrier payload)
        self.peer_counts = None
        # ---- flags ----
        self.flags: Optional[IpcBuffer] = None
        self.peer_flags = None
        self.epoch = [0] * FLAG_SLOTS
        self._nccl_flag = None
        # ---- decode path ----
        self.dy: Optional[IpcBuffer] = None            # y of the decode-sized path, for 64 tokens
        self.dx: Optional[IpcBuffer] = None            # gathered token rows x_all (and their padding masks) for 64 tokens
        self.dpeer_y = self.dpeer_x = None
        self.dws_by_T: Dict[int, EpWorkspace] = {}
        # ---- weight-gather path ----
        self.stage = None                              # [(w13_full, w2_full)] x 2 slots
        self.slot_ready = self.slot_free = None
        self.copy_stream = None
        self.n_fetch = 0

    # ------------------------------------------------------------------ handle exchange (collective, one-off)
    def _exchange(self, bufs: Dict[str, IpcBuffer]) -> Dict[str, List[int]]:
        """All-gather the cudaIpc handles of ``bufs`` over the group and map the peers' buffers.  Collective."""
        import torch.distributed as dist

        lib = _lib.load()
        mine = {k: b.export() for k, b in bufs.items()}
        allh = [None] * self.world
        dist.all_gather_object(allh, mine, group=self.group)
        ptrs = {k: [] for k in bufs}
        with ops.on_device(self.device):
            for r in range(self.world):
                for name in bufs:
                    if r == self.rank:
                        ptrs[name].append(bufs[name].ptr)
                    else:
                        p = ctypes.c_void_p()
                        hb = (ctypes.c_uint8 * 64).from_buffer_copy(allh[r][name])
                        _lib.check(lib.dcmoe_ipc_import(hb, ctypes.byref(p)), "dcmoe_ipc_import")
                        self._imported.append(p.value)
                        ptrs[name].append(p.value)
        return ptrs

    def _ensure_flags(self):
        if self.flags is not None or not self.real:
            return
        self.flags = IpcBuffer(FLAG_SLOTS * MAX_RANKS * 4, self.device, True, zero=True)
        self.counts_table = IpcBuffer(2 * MAX_RANKS * 16 * 4, self.device, True, zero=True)   # two call parities
        torch.cuda.synchronize(self.device)
        ptrs = self._exchange({"flags": self.flags, "counts": self.counts_table})
        self.peer_flags = _ptr_array(ptrs["flags"])
        self._peer_counts_raw = ptrs["counts"]
        self._nccl_flag = torch.zeros(1, dtype=torch.int32, device=self.device)

    def barrier(self, slot: int, payload: Optional[torch.Tensor] = None, peer_dst=None):
        """Stream-ordered barrier of the group on the current stream; with ``payload`` also an all-gather of it into the
        buffers ``peer_dst`` points to (rank q's payload at offset q * payload bytes)."""
        import torch.distributed as dist

        if not self.real:
            return
        if not self.use_flags:
            assert payload is None
            dist.all_reduce(self._nccl_flag, group=self.group)
            return
        self.epoch[slot] += 1
        nbytes = 0 if payload is None else payload.numel() * payload.element_size()
        _lib.check(_lib.load().dcmoe_ep_barrier(self.peer_flags, self.rank, self.world, slot, self.epoch[slot],
                                                None if payload is None else payload.data_ptr(), nbytes, peer_dst,
                                                torch.cuda.current_stream(self.device).cuda_stream), "dcmoe_ep_barrier")

    # ------------------------------------------------------------------ dispatch path
    def row_capacity_for(self, T: int, T_global_max: int) -> int:
        t_pad = (T + 127) // 128 * 128
        return t_pad + self.n_loc * T_global_max + 128 * self.n_loc

    def configure_dispatch(self, max_tokens: int):
        """Allocate (once) the peer-visible dispatch workspace for up to ``max_tokens`` tokens PER RANK and exchange its
        handles.  Collective.  Worst case: every token of every rank selects every expert of one owner."""
        import torch.distributed as dist

        if self.ws_full is not None:
            # growing means freeing buffers peers may still be reading: everyone arrives here first
            torch.cuda.synchronize(self.device)
            if self.real:
                dist.barrier(group=self.group)
            self.close_dispatch()
        self.max_tokens = int(max_tokens)
        cap = self.row_capacity_for(self.max_tokens, self.max_tokens * self.world)
        self.ws_full = EpWorkspace(self.dims, self.dtype, self.max_tokens, self.device, cap, self.real)
        self.ws_by_T = {self.max_tokens: self.ws_full}
        if self.real:
            self._ensure_flags()
            torch.cuda.synchronize(self.device)
            ptrs = self._exchange(self.ws_full.buffers())
            self.peer = (_ptr_array(ptrs["x_packed"]), _ptr_array(ptrs["row_scale"]), _ptr_array(ptrs["y"]))

    def dispatch_workspace(self, T: int) -> EpWorkspace:
        """Workspace of a T-token dispatch call: plan / h / maps sized for T, peer-visible buffers = views of the
        max_tokens workspace (same base addresses on every rank whatever T is)."""
        ws = self.ws_by_T.get(T)
        if ws is None:
            if len(self.ws_by_T) >= 6:
                for k in list(self.ws_by_T):
                    if k != self.max_tokens:
                        self.ws_by_T.pop(k)
                        break
            cap = self.row_capacity_for(T, self.max_tokens * self.world)
            ws = EpWorkspace(self.dims, self.dtype, T, self.device, cap, self.real, shared=self.ws_full.buffers()
                             if self.real else None)
            self.ws_by_T[T] = ws
        return ws

    def close_dispatch(self):
        for ws in self.ws_by_T.values():
            ws.free()
        self.ws_by_T, self.ws_full, self.peer = {}, None, None

    # ------------------------------------------------------------------ weight-gather path
    def ensure_staging(self):
        if self.stage is not None:
            return
        d = self.dims
        G = d.n_real + 1
        mk = lambda: (torch.empty((G, 2 * d.dynamic_intermediate_size, d.hidden_size), dtype=self.dtype, device=self.device),  # noqa: E731
                      torch.empty((G, d.hidden_size, d.dynamic_intermediate_size), dtype=self.dtype, device=self.device))
        self.stage = [mk(), mk()]
        self.slot_ready = [torch.cuda.Event() for _ in range(2)]
        self.slot_free = [torch.cuda.Event() for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(self.device)

    def close(self):
        """Unmap the peers' buffers and free this rank's.  Call on every rank after a barrier (nobody may still be
        reading); process exit does the same implicitly."""
        lib = _lib.load()
        for p in self._imported:
            try:
                lib.dcmoe_ipc_close(p)
            except Exception:  # noqa: BLE001
                pass
        self._imported = []
        self.close_dispatch()
        for b in (self.flags, self.counts_table, self.dy, self.dx):
            if b is not None:
                b.free()
        self.flags = self.counts_table = self.dy = self.dx = None
        for k, v in list(EpContext._registry.items()):
            if v is self:
                EpContext._registry.pop(k)


class ExpertParallelDCMoE:
    """Expert-parallel wrapper around a :class:`DCMoE` that holds (at least) this rank's weights.

    ``group`` is a torch.distributed process group (NCCL) of up to 8 ranks on one node; ``rank`` / ``world``
    default to the group's.  The call signature and the returned 6-tuple are those of the reference block."""

    def __init__(self, module: DCMoE, group=None, rank: Optional[int] = None, world: Optional[int] = None):
        import torch.distributed as dist

        self.m = module
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        d = module.dims
        if d.n_real % self.world != 0 or self.world > MAX_RANKS:
            raise ValueError(f"num_experts ({d.n_real}) must be divisible by ep_size ({self.world}) (core.py:505), ep_size <= 8")
        self.n_loc = d.n_real // self.world
        self.local_dims = replace(d, n_real=self.n_loc)
        self._w13 = self._w2 = None
        self._wbuf: Optional[Dict[str, IpcBuffer]] = None
        self._peer_w = None                # (w13, w2) pointer arrays over the ranks (weight-gather path)
        self.ctx: Optional[EpContext] = None
        self.ws: Optional[EpWorkspace] = None
        self.mode = os.environ.get("DCMOE_EP_MODE", "auto")                      # auto | dispatch | gather
        self.gather_min_tokens = int(os.environ.get("DCMOE_EP_GATHER_MIN_TOKENS", "8192"))
        self.check_lockstep = os.environ.get("DCMOE_EP_CHECK", "0") == "1"       # debug: host all-gather of T per call
        self.overlap = os.environ.get("DCMOE_EP_OVERLAP", "1") != "0"   # dispatch path: comm stream under the shared experts' GEMMs
        self.comm_ctas = int(os.environ.get("DCMOE_EP_COMM_CTAS", "148"))   # grid cap of dispatch / partial combine (0 = full)
        self.gemm_ctas = int(os.environ.get("DCMOE_EP_GEMM_CTAS", "0"))   # CTAs of the shared GEMMs that run under comm (0 = all SMs)
        self._side = None
        self.comm_events = []        # (name, start, end) CUDA events of the comm / copy streams, filled when a stage hook is set
        self._local_cfg = None
        # decode-sized calls (world * T <= 64): "replicate" (default) = keep a resident copy of the remote experts' packs
        # (fetched once over NVLink, 270 MB x (R-1)/R per layer) and run the call with no per-call exchange at all;
        # "exchange" = replicate the TOKENS instead (decode_forward); "0" = take the large-T paths.  Measured at 8 GPUs
        # (profiles/r02_ep_decode_latency.txt): exchange 98 - 107 us per layer call against 70 - 82 us on one GPU (and 71 - 79 us for the replica) -- at this size the
        # layer is bound by per-kernel fixed costs, not by the weight bytes a rank streams, so sharding them buys nothing
        env_d = os.environ.get("DCMOE_EP_DECODE", "replicate")
        self.decode_policy = {"1": "exchange", "0": "off"}.get(env_d, env_d)
        self.decode_mode = self.decode_policy != "off"
        self._resident = None              # (w13_full, w2_full): all experts' packs, kept (decode policy "replicate")
        self._dws: Optional[EpWorkspace] = None
        self.last_path = None
        self._call = 0

    @property
    def kernels_per_step(self) -> int:
        """Kernels of this package launched per forward on the path taken last (bench.py's gpu_launches)."""
        return {"gather": 6, "decode": 6, "dispatch": 14 if self.overlap else 11}.get(self.last_path or "dispatch", 11)

    def context(self, dtype, device) -> EpContext:
        if self.ctx is None or self.ctx.dtype != dtype or self.ctx.device != torch.device(device):
            self.ctx = EpContext.get(self.group, self.rank, self.world, self.m.dims, dtype, device)
        return self.ctx

    # ------------------------------------------------------------------ weights
    def _alloc_local_packs(self, dtype, device):
        d = self.m.dims
        G = self.n_loc + 1
        es = torch.empty((), dtype=dtype).element_size()
        n13 = G * 2 * d.dynamic_intermediate_size * d.hidden_size * es
        n2 = G * d.hidden_size * d.dynamic_intermediate_size * es
        ipc = self.group is not None
        self._wbuf = {"w13": IpcBuffer(n13, device, ipc), "w2": IpcBuffer(n2, device, ipc)}
        self._w13 = self._wbuf["w13"].view(dtype, (G, 2 * d.dynamic_intermediate_size, d.hidden_size))
        self._w2 = self._wbuf["w2"].view(dtype, (G, d.hidden_size, d.dynamic_intermediate_size))

    def pack_local_weights(self):
        """This rank's packs: its n_loc routed experts as groups 0..n_loc-1, the shared pair as group n_loc.  They live in
        cudaIpc-exportable memory so that the peers can pull them (weight-gather path)."""
        if self._w13 is not None:
            return
        m, d, ld = self.m, self.m.dims, self.local_dims
        p = m.gate.weight
        self._alloc_local_packs(p.dtype, p.device)
        routed, shared = m._expert_params()
        for l in range(self.n_loc):
            e = self.rank * self.n_loc + l
            ex = routed[e] if len(routed) == d.n_real else routed[l]   # full module or local-experts-only module
            ops.pack_expert(ex.gate_proj.weight.detach().contiguous(), ex.up_proj.weight.detach().contiguous(),
                            ex.down_proj.weight.detach().contiguous(), l, 0, ld, self._w13, self._w2)
        for i, ex in enumerate(shared):
            ops.pack_expert(ex.gate_proj.weight.detach().contiguous(), ex.up_proj.weight.detach().contiguous(),
                            ex.down_proj.weight.detach().contiguous(), self.n_loc, i, ld, self._w13, self._w2)

    def set_packed_local_weights(self, w13: torch.Tensor, w2: torch.Tensor):
        """Use packs built elsewhere (``checkpoint.load_dcmoe_ep``: this rank's experts read straight from a
        checkpoint) instead of packing from the wrapped module's parameters."""
        d, G = self.m.dims, self.n_loc + 1
        if tuple(w13.shape) != (G, 2 * d.dynamic_intermediate_size, d.hidden_size) or \
                tuple(w2.shape) != (G, d.hidden_size, d.dynamic_intermediate_size) or not (w13.is_cuda and w2.is_cuda):
            raise ValueError("packed local weights have the wrong shape for this rank's expert count")
        if self.group is None:       # virtual ranks (one process): the tensors are used as they are
            self._w13, self._w2 = w13.contiguous(), w2.contiguous()
            self._wbuf = {}
            for name, t in (("w13", self._w13), ("w2", self._w2)):
                b = IpcBuffer.__new__(IpcBuffer)
                b.nbytes, b.device, b.ipc, b._torch, b.ptr = t.numel() * t.element_size(), t.device, False, t, t.data_ptr()
                self._wbuf[name] = b
        else:                        # separate processes: the packs must live in cudaIpc-exportable memory
            self._alloc_local_packs(w13.dtype, w13.device)
            self._w13.copy_(w13)
            self._w2.copy_(w2)
        self._peer_w = None

    def set_peer_weights(self, w13: List[int], w2: List[int]):
        self._peer_w = (_ptr_array(w13), _ptr_array(w2))

    def _exchange_weight_handles(self):
        """Collective, once per layer: map every rank's packs.  The barrier makes sure they are packed."""
        import torch.distributed as dist

        ctx = self.ctx
        torch.cuda.synchronize(ctx.device)
        ptrs = ctx._exchange(self._wbuf)
        dist.barrier(group=self.group)
        self.set_peer_weights(ptrs["w13"], ptrs["w2"])

    # ------------------------------------------------------------------ dispatch path: phases (all launch-only)
    def phase_route(self, x: torch.Tensor, attention_mask=None, router_logits=None):
        m, ws = self.m, self.ws
        hook = m.stage_hook or (lambda _n: None)
        hook("start")
        wg = m.gate.weight.detach()
        self._x = x
        if router_logits is None:
            router_logits = getattr(self, "_forced_logits", None)     # tests: identical logits on every path
        self._route = ops.router(x, wg, ws, logits_in=router_logits, attention_mask=attention_mask)
        hook("router")
        ops.plan(ws)                     # local counts + block prefix sums (+ local aux loss)
        self._aux = ws.aux_loss.clone().reshape(())
        if getattr(ws, "_t_const", None) is None:      # cached: a fresh torch.tensor(..., device=) is a blocking H2D copy
            ws._t_const = torch.tensor([ws.T], dtype=torch.int32, device=ws.device)
        vec = torch.cat([ws.counts, ws._t_const])
        hook("plan")
        return vec

    def phase_plan(self, all_counts: torch.Tensor):
        """ep_plan: row-space layout of this rank + destinations of its rows (+ shared-expert row scales)."""
        lib = _lib.load()
        ws = self.ws
        hook = self.m.stage_hook or (lambda _n: None)
        st = torch.cuda.current_stream().cuda_stream
        self._all_counts = all_counts
        _, _, _mask, gw = self._route
        _lib.check(lib.dcmoe_ep_plan(self._all_counts.data_ptr(), self.rank, self.world, ws.T, ws.row_capacity, ws.cfg,
                                     ws.plan.data_ptr(), ws.ep_meta.data_ptr(), gw.data_ptr(), ws.row_scale.data_ptr(), st),
                   "dcmoe_ep_plan")
        hook("ep_plan")

    def phase_dispatch(self, max_ctas: int = 0):
        lib = _lib.load()
        ws = self.ws
        st = torch.cuda.current_stream().cuda_stream
        _, _, mask, gw = self._route
        xp, rs, _y = self._peer
        _lib.check(lib.dcmoe_ep_dispatch(self._x.data_ptr(), mask.data_ptr(), gw.data_ptr(), ws.T, ws.row_capacity, ws.cfg,
                                         ws.plan.data_ptr(), ws.ep_meta.data_ptr(), self.rank, self.world, xp, rs,
                                         ws.slot_of.data_ptr(), max_ctas, st), "dcmoe_ep_dispatch")

    def phase_ffn(self, phase: int = 0, group_sel: int = 0, name: Optional[str] = None, max_ctas: int = 0):
        """phase 0/1/2 = both / GEMM-1 / GEMM-2; group_sel 0/1/2 = all / shared-expert / routed row tiles."""
        lib = _lib.load()
        ws = self.ws
        if self._local_cfg is None:
            self._local_cfg = self.local_dims.c_config(ws.dtype)
        impl = self.m.ffn_impl if self.m.ffn_impl is not None else (0 if ws.dtype == torch.bfloat16 else 1)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.dcmoe_grouped_ffn(self._x.data_ptr(), ws.x_packed.data_ptr(), self._w13.data_ptr(),
                                         self._w2.data_ptr(), ws.row_scale.data_ptr(), ws.T, ws.row_capacity, self._local_cfg,
                                         ws.plan.data_ptr(), ws.h.data_ptr(), ws.y.data_ptr(), impl,
                                         phase | (group_sel << 4) | (max_ctas << 8) | (1 << 20), st),
                   "dcmoe_grouped_ffn")
        if name and self.m.stage_hook:
            self.m.stage_hook(name)

    def phase_combine(self, out: Optional[torch.Tensor], mode: int = 0, max_ctas: int = 0):
        lib = _lib.load()
        ws = self.ws
        _xp, _rs, y = self._peer
        _lib.check(lib.dcmoe_ep_combine(ws.y.data_ptr(), y, ws.slot_of.data_ptr(), ws.T, ws.cfg, self.world, mode,
                                        ws.partial.data_ptr() if mode != 0 else None,
                                        None if out is None else out.data_ptr(), max_ctas,
                                        torch.cuda.current_stream().cuda_stream), "dcmoe_ep_combine")

    def set_peers(self, x_packed: List[int], row_scale: List[int], y: List[int]):
        self._peer = (_ptr_array(x_packed), _ptr_array(row_scale), _ptr_array(y))

    # ------------------------------------------------------------------ decode-sized calls
    def decode_applicable(self, T: int, dtype) -> bool:
        return (self.decode_mode and dtype == torch.bfloat16 and 0 < T * self.world <= 64 and self.m.use_front_small
                and self.m.ffn_impl in (None, 0, 3))

    def ensure_decode_workspace(self, T_total: int, device, ipc: bool) -> EpWorkspace:
        """Workspace of a decode-sized call on T_total gathered tokens (kept per token count).  Across processes the
        peer-visible buffers (y, and x_all which the peers fill) are allocated ONCE, for 64 tokens, and every per-T
        workspace views them, so the cudaIpc handles are exchanged a single time however the batch size varies."""
        ctx = self.context(torch.bfloat16, device)
        ws = ctx.dws_by_T.get(T_total)
        if ws is None:
            H = self.m.dims.hidden_size
            if len(ctx.dws_by_T) >= 8:
                ctx.dws_by_T.pop(next(iter(ctx.dws_by_T)))
            if ipc and ctx.dy is None:
                sizes, _ = ops.query_sizes(self.m.dims, torch.bfloat16, 64, 0)
                ctx.dy = IpcBuffer(int(sizes.row_capacity) * H * 2, device, True)
                ctx.dx = IpcBuffer(64 * H * 2 + 64 * 4, device, True, zero=True)   # rows, then int32 padding masks
                ctx._ensure_flags()
                torch.cuda.synchronize(ctx.device)
                ptrs = ctx._exchange({"dy": ctx.dy, "dx": ctx.dx})
                ctx.dpeer_y = _ptr_array(ptrs["dy"])
                ctx.dpeer_x = _ptr_array(ptrs["dx"])
                ctx.dpeer_am = _ptr_array([p + 64 * H * 2 for p in ptrs["dx"]])
            ws = EpWorkspace(self.m.dims, torch.bfloat16, T_total, device, 0, False)   # worst-case rows for T_total tokens
            if ipc:
                ws.y = _wrap(ctx.dy.ptr, ws.row_capacity * H * 2, torch.bfloat16, (ws.row_capacity, H), ws.device)
                ws.x_all = ctx.dx.view(torch.uint8, (ctx.dx.nbytes,))[: T_total * H * 2].view(torch.bfloat16).view(T_total, H)
                ws.am_all = ctx.dx.view(torch.uint8, (ctx.dx.nbytes,))[64 * H * 2: 64 * H * 2 + T_total * 4].view(torch.int32)
                ws._c_ws = None
            else:
                ws.x_all = torch.empty((T_total, H), dtype=torch.bfloat16, device=device)
                ws.am_all = torch.ones(T_total, dtype=torch.int32, device=device)
            ctx.dws_by_T[T_total] = ws
        self._dws = ws
        return ws

    def set_decode_peers(self, y: List[int]):
        self._dpeer = _ptr_array(y)

    def decode_route_and_ffn(self, attention_mask_all=None):
        """Replicated part of a decode-sized call, after x_all holds every rank's tokens (rank-major): fused front end
        on all tokens (identical plan and row space on every rank), then the weight-streaming GEMMs over the row tiles
        whose weights this rank holds (its routed experts + the shared pair)."""
        lib = _lib.load()
        ws = self._dws
        self._droute = ops.front_small(ws.x_all, self.m.gate.weight.detach(), ws, attention_mask=attention_mask_all)
        _lib.check(lib.dcmoe_grouped_ffn(ws.x_all.data_ptr(), ws.x_packed.data_ptr(), self._w13.data_ptr(), self._w2.data_ptr(),
                                         ws.row_scale.data_ptr(), ws.T, ws.row_capacity, ws.cfg, ws.plan.data_ptr(),
                                         ws.h.data_ptr(), ws.y.data_ptr(), 3, (self.n_loc << 21) | (self.rank << 25),
                                         torch.cuda.current_stream().cuda_stream), "dcmoe_grouped_ffn")

    def decode_combine(self, T: int, out: torch.Tensor):
        """Own tokens [rank * T, (rank + 1) * T): routed rows from the owners' y (peer loads), shared row from the local y."""
        lib = _lib.load()
        ws, d = self._dws, self.m.dims
        es = 2
        off = self.rank * T
        _lib.check(lib.dcmoe_ep_combine(ws.y.data_ptr() + off * d.hidden_size * es, self._dpeer,
                                        ws.slot_of.data_ptr() + off * d.n_real * 4, T, ws.cfg, self.world, 0, None,
                                        out.data_ptr(), 0, torch.cuda.current_stream().cuda_stream), "dcmoe_ep_combine")

    def decode_forward(self, hidden_states: torch.Tensor, attention_mask=None):
        """Expert parallelism for decode-sized calls.  The dispatch / combine exchange of the large-T path costs a dozen
        launches; with a handful of tokens it is cheaper to replicate the tokens (every rank pushes its rows into every
        rank's x_all: one barrier-with-payload launch) and the routing, let every rank stream only ITS experts' weights,
        and gather the routed rows in the combine.  Same kernels and row space as a single-GPU call on the gathered
        tokens, so the output equals that call's bit for bit.  (aux_loss is then the loss over the gathered tokens.)"""
        import torch.distributed as dist

        B, S, H = hidden_states.shape
        T = B * S
        x = hidden_states.reshape(T, H)
        if not x.is_contiguous():
            x = x.contiguous()
        self.pack_local_weights()
        ws = self.ensure_decode_workspace(T * self.world, x.device, ipc=True)
        ctx = self.ctx
        self.set_decode_peers(list(ctx.dpeer_y))
        am_all = None
        if ctx.use_flags:
            # "every rank is done reading the previous call's y / x_all" is implied: a rank only passes the y barrier
            # of call i after every rank arrived there, i.e. after every rank's combine of call i-1 ... and the x push
            # of call i+1 can only overwrite rows whose readers (front end of call i) precede that rank's y barrier
            if attention_mask is not None:
                am = attention_mask.reshape(-1).to(device=x.device, dtype=torch.int32).contiguous()
                ctx.barrier(SLOT_DECODE_X, am, ctx.dpeer_am)
                am_all = ws.am_all
            ctx.barrier(SLOT_DECODE_X, x, ctx.dpeer_x)
        else:
            dist.all_gather_into_tensor(ws.x_all, x, group=self.group)
            if attention_mask is not None:
                am = attention_mask.reshape(-1).to(device=x.device, dtype=torch.int32).contiguous()
                am_all = torch.empty(T * self.world, dtype=torch.int32, device=x.device)
                dist.all_gather_into_tensor(am_all, am, group=self.group)
        self.decode_route_and_ffn(am_all)
        ctx.barrier(SLOT_DECODE_Y)                                     # every owner's y rows are complete
        out = torch.empty((B, S, H), dtype=x.dtype, device=x.device)
        self.decode_combine(T, out)
        logits, top_k, mask, gw = (t[self.rank * T:(self.rank + 1) * T] for t in self._droute)
        self.m.last_workspace = ws
        return out, logits, top_k, mask, gw, ws.aux_loss.clone().reshape(())

    # ------------------------------------------------------------------ weight-gather path
    def fetch_weights(self, slot: int):
        """Enqueue, on the context's copy stream, the copies of every rank's packs into staging slot ``slot``."""
        ctx = self.ctx
        lib = _lib.load()
        w13_full, w2_full = ctx.stage[slot]
        cs = ctx.copy_stream
        timed = self.m.stage_hook is not None
        with torch.cuda.stream(cs):
            cs.wait_event(ctx.slot_free[slot])          # the GEMMs that last read this slot are done
            if timed:
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record(cs)
            _lib.check(lib.dcmoe_ep_fetch_weights(self._peer_w[0], self._peer_w[1], self.rank, self.world,
                                                  self.m.dims.c_config(ctx.dtype), w13_full.data_ptr(), w2_full.data_ptr(),
                                                  cs.cuda_stream), "dcmoe_ep_fetch_weights")
            if timed:
                t1.record(cs)
                self.comm_events.append(("weight_fetch", t0, t1))
            ctx.slot_ready[slot].record(cs)

    def gather_forward(self, hidden_states: torch.Tensor, attention_mask=None, aux_balance_weight=None, router_logits=None):
        """Weight-gather expert parallelism: pull every remote expert's packed weights over NVLink (copy engines) into a
        staging pack, then run the single-GPU forward on this rank's tokens -- router, plan and permute run while the
        copies are in flight; the staging packs are double buffered, so with the host running ahead the fetch of call
        i+1 overlaps the GEMMs of call i.  Bit-equal to the single-GPU layer on the same rows (the FFN is row
        independent); the aux loss is the loss over this rank's tokens, as on the dispatch path."""
        ctx = self.context(hidden_states.dtype, hidden_states.device)
        self.pack_local_weights()
        if self._peer_w is None:
            self._exchange_weight_handles()
        ctx.ensure_staging()
        slot = ctx.n_fetch & 1
        ctx.n_fetch += 1
        main = torch.cuda.current_stream(ctx.device)
        self.fetch_weights(slot)
        w13_full, w2_full = ctx.stage[slot]
        out = self.m._forward_local(hidden_states, attention_mask, aux_balance_weight, router_logits, None, w13_full, w2_full,
                                    before_ffn=lambda: main.wait_event(ctx.slot_ready[slot]))
        ctx.slot_free[slot].record(main)
        return out

    def resident_forward(self, hidden_states: torch.Tensor, attention_mask=None, aux_balance_weight=None, router_logits=None):
        """Decode-sized calls with the experts replicated: the first call pulls every remote expert's packed weights
        (collective once per layer: the handle exchange) into a pack that STAYS resident; every later call is the
        single-GPU forward on this rank's tokens -- no barrier, no exchange, CUDA-graph capturable like the plain layer.
        Inference weights are static; call ``drop_resident_weights()`` after changing them."""
        ctx = self.context(hidden_states.dtype, hidden_states.device)
        if self._resident is None:
            self.pack_local_weights()
            if self._peer_w is None:
                self._exchange_weight_handles()
            d = self.m.dims
            G = d.n_real + 1
            w13 = torch.empty((G, 2 * d.dynamic_intermediate_size, d.hidden_size), dtype=ctx.dtype, device=ctx.device)
            w2 = torch.empty((G, d.hidden_size, d.dynamic_intermediate_size), dtype=ctx.dtype, device=ctx.device)
            _lib.check(_lib.load().dcmoe_ep_fetch_weights(self._peer_w[0], self._peer_w[1], self.rank, self.world,
                                                          d.c_config(ctx.dtype), w13.data_ptr(), w2.data_ptr(),
                                                          torch.cuda.current_stream(ctx.device).cuda_stream),
                       "dcmoe_ep_fetch_weights")
            self._resident = (w13, w2)
        return self.m._forward_local(hidden_states, attention_mask, aux_balance_weight, router_logits, None, *self._resident)

    def drop_resident_weights(self):
        self._resident = None

    # ------------------------------------------------------------------ distributed forward
    def _check_lockstep(self, T: int, path: str):
        import torch.distributed as dist

        got = [None] * self.world
        dist.all_gather_object(got, (T, path), group=self.group)
        if len({p for _t, p in got}) != 1 or (path == "decode" and len({t for t, _p in got}) != 1):
            raise RuntimeError(f"expert-parallel ranks diverged: (token count, path) per rank = {got}; every rank must "
                               f"call with token counts on the same side of the decode (world*T <= 64) and weight-gather "
                               f"(T >= {self.gather_min_tokens}) thresholds, and with equal T on the decode path")
        return got

    @torch.no_grad()
    def __call__(self, hidden_states: torch.Tensor, attention_mask=None, aux_balance_weight=None, router_logits=None):
        with ops.on_device(hidden_states.device):
            return self._call_impl(hidden_states, attention_mask, aux_balance_weight, router_logits)

    def _call_impl(self, hidden_states, attention_mask, aux_balance_weight, router_logits):
        import torch.distributed as dist

        B, S, H = hidden_states.shape
        T = B * S
        path = choose_path(T, self.world, hidden_states.dtype, self.mode, self.gather_min_tokens,
                           decode_ok=(router_logits is None and self.m.stage_hook is None and aux_balance_weight is None
                                      and self.decode_applicable(T, hidden_states.dtype)))
        if self.m.token_drop or self.m.training or aux_balance_weight is not None:
            # the capacity branch, the training-mode gate and the weighted aux loss are per-rank computations on the
            # rank's own tokens (core.py:293-329 runs before the exchange): they ride on the weight-gather path, which is
            # the single-GPU forward over local tokens; the token-dispatch kernels do not carry them
            if self.mode == "dispatch":
                raise NotImplementedError("token_drop / training-mode forward / aux_balance_weight need the weight-gather "
                                          "expert-parallel path (DCMOE_EP_MODE=auto or gather)")
            path = "gather"
        if self.check_lockstep:
            self._check_lockstep(T, path)
        self.last_path = path
        self._call += 1
        if path == "decode" and self.decode_policy == "replicate":
            out = self.resident_forward(hidden_states, attention_mask)
        elif path == "decode":
            out = self.decode_forward(hidden_states, attention_mask)
        elif path == "gather":
            out = self.gather_forward(hidden_states, attention_mask, aux_balance_weight, router_logits)
        else:
            out = self.dispatch_forward(hidden_states, attention_mask, router_logits)
        if getattr(self.m, "avg_hidden_states_last", False):
            # core.py:355-356: all_reduce(final_hidden_states, AVG) over the expert-parallel group in eval mode
            dist.all_reduce(out[0], op=dist.ReduceOp.SUM, group=self.group)
            out[0].div_(self.world)
        return out

    def dispatch_forward(self, hidden_states: torch.Tensor, attention_mask=None, router_logits=None):
        import torch.distributed as dist

        B, S, H = hidden_states.shape
        T = B * S
        x = hidden_states.reshape(T, H)
        if not x.is_contiguous():
            x = x.contiguous()
        self.pack_local_weights()
        ctx = self.context(x.dtype, x.device)
        if ctx.ws_full is None:
            # first dispatch call of the group (collective by construction): agree on the per-rank token bound.  The
            # workspace serves every later call with T <= max_tokens on every rank; larger calls take the weight-gather
            # path in auto mode, so it never has to grow there.
            got = [None] * self.world
            dist.all_gather_object(got, T, group=self.group)
            bound = max(got) if self.mode == "dispatch" else max(max(got), self.gather_min_tokens)
            ctx.configure_dispatch(bound)
        if T > ctx.max_tokens:
            raise RuntimeError(f"expert-parallel dispatch workspace holds {ctx.max_tokens} tokens per rank, got {T}: call "
                               f"ExpertParallelDCMoE.ctx.configure_dispatch(max_tokens) on every rank first")
        self.ws = ctx.dispatch_workspace(T)
        self._peer = ctx.peer
        self.m.last_workspace = self.ws
        hook = self.m.stage_hook or (lambda _n: None)
        main = torch.cuda.current_stream()
        vec = self.phase_route(x, attention_mask, router_logits)
        n = vec.numel()
        if ctx.use_flags:
            # all-gather of (counts, T) through peer memory; also the "buffers are free" barrier: a rank gets here only
            # after its combine of the previous call, so when everyone has arrived nobody still reads anybody's y
            par = ctx.epoch[SLOT_COUNTS] & 1          # two tables, alternating from one exchange to the next
            dst = _ptr_array([p + par * MAX_RANKS * 16 * 4 for p in ctx._peer_counts_raw])
            ctx.barrier(SLOT_COUNTS, vec, dst)
            all_counts = ctx.counts_table.view(torch.int32, (2 * MAX_RANKS * 16,))[par * MAX_RANKS * 16:
                                                                                    par * MAX_RANKS * 16 + self.world * n].view(self.world, n)
        else:
            all_counts = torch.empty((self.world, n), dtype=torch.int32, device=x.device)
            dist.all_gather_into_tensor(all_counts, vec, group=self.group)
        hook("allgather_counts")
        self.phase_plan(all_counts)
        out = torch.empty((B, S, H), dtype=x.dtype, device=x.device)
        tcgen05 = (self.m.ffn_impl in (None, 0)) and x.dtype == torch.bfloat16
        if not (self.overlap and tcgen05):
            self.phase_dispatch()
            hook("ep_dispatch")
            ctx.barrier(SLOT_DISPATCH)                                            # every rank's rows have landed
            hook("barrier_dispatch")
            self.phase_ffn(1, 0, "ffn_gemm1")
            self.phase_ffn(2, 0, "ffn_gemm2")
            ctx.barrier(SLOT_FFN)                                                 # every owner's y is complete
            hook("barrier_ffn")
            self.phase_combine(out.view(T, H), 0)
            hook("ep_combine")
        else:
            # comm stream: dispatch over NVLink + barrier, while the main stream runs the shared experts' GEMM-1
            if self._side is None:
                self._side = torch.cuda.Stream(x.device)
                self._ev = [torch.cuda.Event() for _ in range(4)]
            side, ev = self._side, self._ev
            ev[0].record(main)
            timed = self.m.stage_hook is not None
            with torch.cuda.stream(side):
                side.wait_event(ev[0])
                if timed:
                    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    t0.record(side)
                self.phase_dispatch(self.comm_ctas)
                if timed:
                    t1.record(side)
                    self.comm_events.append(("ep_dispatch", t0, t1))
                ctx.barrier(SLOT_DISPATCH)                                        # barrier 1 (on the comm stream)
                ev[1].record(side)
            self.phase_ffn(1, 1, "ffn_gemm1_shared", self.gemm_ctas)             # overlaps the dispatch
            main.wait_event(ev[1])
            hook("wait_dispatch")
            self.phase_ffn(1, 2, "ffn_gemm1_routed")
            self.phase_ffn(2, 2, "ffn_gemm2_routed")
            ctx.barrier(SLOT_FFN)                                                 # barrier 2: routed y complete everywhere
            hook("barrier_ffn")
            ev[2].record(main)
            with torch.cuda.stream(side):
                side.wait_event(ev[2])
                if timed:
                    t2, t3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    t2.record(side)
                self.phase_combine(None, 1, self.comm_ctas)                      # routed rows over NVLink -> fp32 partial
                if timed:
                    t3.record(side)
                    self.comm_events.append(("ep_combine_gather", t2, t3))
                ev[3].record(side)
            self.phase_ffn(2, 1, "ffn_gemm2_shared", self.gemm_ctas)
            main.wait_event(ev[3])
            hook("wait_combine")
            self.phase_combine(out.view(T, H), 2)
            hook("ep_combine_final")
        logits, top_k, mask, gw = self._route
        return out, logits, top_k, mask, gw, self._aux


class LocalRanks:
    """R virtual ranks in one process on one GPU: the same kernels and call order as the distributed forward,
    with the collectives replaced by torch ops on the ranks' tensors.  Test / debugging aid."""

    def __init__(self, module: DCMoE, world: int, split: bool = False):
        self.world = world
        self.split = split
        self.ranks = [ExpertParallelDCMoE(module, group=None, rank=r, world=world) for r in range(world)]

    @torch.no_grad()
    def decode_forward(self, xs: Sequence[torch.Tensor], attention_masks=None):
        """The decode-sized expert-parallel path (ExpertParallelDCMoE.decode_forward) with the push of the token rows
        replaced by a torch.cat and the peers' y buffers by the virtual ranks' own.  Every x must have the same token count."""
        flat = [x.reshape(-1, x.shape[-1]).contiguous() for x in xs]
        T = flat[0].shape[0]
        assert all(f.shape[0] == T for f in flat)
        x_all = torch.cat(flat)
        am_all = None if attention_masks is None else torch.cat([a.reshape(-1).to(torch.int32) for a in attention_masks])
        for ep in self.ranks:
            ep.pack_local_weights()
            ws = ep.ensure_decode_workspace(T * self.world, x_all.device, ipc=False)
            ws.x_all.copy_(x_all)
        for ep in self.ranks:
            ep.set_decode_peers([q._dws.y.data_ptr() for q in self.ranks])
            ep.decode_route_and_ffn(am_all)
        outs = []
        for r, ep in enumerate(self.ranks):
            out = torch.empty_like(flat[r])
            ep.decode_combine(T, out)
            logits, top_k, mask, gw = (t[r * T:(r + 1) * T] for t in ep._droute)
            outs.append((out.view(xs[r].shape), logits, top_k, mask, gw, ep._dws.aux_loss.clone().reshape(())))
        return outs

    @torch.no_grad()
    def resident_forward(self, xs: Sequence[torch.Tensor], attention_masks=None):
        """Decode policy "replicate": every virtual rank keeps a resident copy of all packs and runs its tokens locally."""
        for ep in self.ranks:
            ep.pack_local_weights()
            ep.context(xs[0].dtype, xs[0].device)
        for ep in self.ranks:
            ep.set_peer_weights([q._wbuf["w13"].ptr for q in self.ranks], [q._wbuf["w2"].ptr for q in self.ranks])
        return [ep.resident_forward(xs[r], None if attention_masks is None else attention_masks[r])
                for r, ep in enumerate(self.ranks)]

    @torch.no_grad()
    def gather_forward(self, xs: Sequence[torch.Tensor], attention_masks=None):
        """The weight-gather path: every virtual rank copies all ranks' packs (local pointers here) into its staging
        pack and runs the single-GPU forward on its own tokens."""
        for ep in self.ranks:
            ep.pack_local_weights()
            ep.context(xs[0].dtype, xs[0].device)
        for ep in self.ranks:
            ep.set_peer_weights([q._wbuf["w13"].ptr for q in self.ranks], [q._wbuf["w2"].ptr for q in self.ranks])
        return [ep.gather_forward(xs[r], None if attention_masks is None else attention_masks[r])
                for r, ep in enumerate(self.ranks)]

    @torch.no_grad()
    def forward(self, xs: Sequence[torch.Tensor], attention_masks=None):
        shapes = [x.shape for x in xs]
        flat = [x.reshape(-1, x.shape[-1]).contiguous() for x in xs]
        T_max = max(f.shape[0] for f in flat)
        for r, ep in enumerate(self.ranks):
            ep.pack_local_weights()
            ctx = ep.context(flat[r].dtype, flat[r].device)
            if ctx.ws_full is None or ctx.max_tokens < T_max:
                ctx.configure_dispatch(T_max)
            ep.ws = ctx.dispatch_workspace(flat[r].shape[0])
            ep.m.last_workspace = ep.ws
        for ep in self.ranks:
            ep.set_peers([q.ws.x_packed.data_ptr() for q in self.ranks], [q.ws.row_scale.data_ptr() for q in self.ranks],
                         [q.ws.y.data_ptr() for q in self.ranks])
        vecs = [ep.phase_route(flat[r], None if attention_masks is None else attention_masks[r]) for r, ep in enumerate(self.ranks)]
        all_counts = torch.stack(vecs).contiguous()
        for ep in self.ranks:
            ep.phase_plan(all_counts)
        if not self.split:
            for ep in self.ranks:
                ep.phase_dispatch()
            for ep in self.ranks:
                ep.phase_ffn(0, 0)
            outs = []
            for r, ep in enumerate(self.ranks):
                out = torch.empty_like(flat[r])
                ep.phase_combine(out, 0)
                logits, top_k, mask, gw = ep._route
                outs.append((out.view(shapes[r]), logits, top_k, mask, gw, ep._aux))
        else:   # the kernel sequence of the overlapped schedule (run serially here)
            for ep in self.ranks:
                ep.phase_ffn(1, 1)          # the shared GEMM-1 needs nothing from the dispatch
            for ep in self.ranks:
                ep.phase_dispatch()
            for ep in self.ranks:
                ep.phase_ffn(1, 2)
                ep.phase_ffn(2, 2)
            for ep in self.ranks:
                ep.phase_combine(None, 1)
            for ep in self.ranks:
                ep.phase_ffn(2, 1)
            outs = []
            for r, ep in enumerate(self.ranks):
                out = torch.empty_like(flat[r])
                ep.phase_combine(out, 2)
                logits, top_k, mask, gw = ep._route
                outs.append((out.view(shapes[r]), logits, top_k, mask, gw, ep._aux))
        self.all_counts = all_counts
        return outs
