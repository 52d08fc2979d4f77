"""Expert-parallel DCMoE (reference: AudioMOELayer.forward with an ep_group, core.py:446-493; group wiring
core.py:505-520).

One process per GPU.  Rank r owns routed experts [r*n_loc, (r+1)*n_loc) (core.py:505), the gate and the shared
experts are replicated, every rank routes its own tokens.  Per forward and per rank:

    router -> local plan -> all-gather of (counts, T)  [NCCL, 36 bytes per rank]
           -> ep_plan -> ep_dispatch (rows stored straight into the owners' packed buffers over NVLink)
           -> barrier -> grouped FFN on the rows this rank owns (tcgen05) -> barrier
           -> ep_combine (routed rows gathered from the owners' y buffers over NVLink + local shared row)

Buffers touched by peers (x_packed, y, row_scale) are allocated with ``dcmoe_ipc_alloc`` and mapped into
every rank with cudaIpc handles exchanged once per workspace.  ``LocalRanks`` runs the same kernels for R
virtual ranks inside one process on one GPU (peer pointers are then ordinary local pointers); it is what the
single-GPU tests use, as separate processes spinning on one GPU is not allowed on the test boxes.
"""
from __future__ import annotations

import ctypes
from dataclasses import replace
from typing import List, Optional, Sequence

import torch

from . import _lib, ops
from .dcmoe import DCMoE
from .ops import LayerDims, Workspace

EP_META_INTS = 32


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


class _RawCuda:
    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _wrap(ptr: int, nbytes: int, dtype, shape, device) -> torch.Tensor:
    return torch.as_tensor(_RawCuda(ptr, nbytes), device=device).view(dtype).view(shape)


class EpWorkspace(Workspace):
    """Workspace whose peer-visible buffers come from dcmoe_ipc_alloc (exportable with cudaIpcGetMemHandle)."""

    def __init__(self, dims: LayerDims, dtype, T: int, device, row_capacity: int, ipc: bool):
        super().__init__(dims, dtype, T, device, row_capacity, alloc_peer_visible=not ipc)
        self.ep_meta = torch.zeros(EP_META_INTS, dtype=torch.int32, device=self.device)
        self.ipc = ipc
        self._raw = {}
        if ipc:
            lib = _lib.load()
            es = torch.empty((), dtype=dtype).element_size()
            with torch.cuda.device(self.device):
                for name, dt, esz in (("x_packed", dtype, es), ("y", dtype, es), ("row_scale", torch.float32, 4)):
                    shape = self.shapes[name]
                    nbytes = shape[0] * shape[1] * esz
                    p = ctypes.c_void_p()
                    _lib.check(lib.dcmoe_ipc_alloc(nbytes, ctypes.byref(p)), "dcmoe_ipc_alloc")
                    self._raw[name] = (p.value, nbytes)
                    setattr(self, name, _wrap(p.value, nbytes, dt, shape, self.device))
            self.row_scale.zero_()

    def export_handles(self) -> dict:
        lib = _lib.load()
        out = {}
        for name, (ptr, _n) in self._raw.items():
            buf = (ctypes.c_uint8 * 64)()
            _lib.check(lib.dcmoe_ipc_export(ptr, buf), "dcmoe_ipc_export")
            out[name] = bytes(buf)
        return out

    def __del__(self):  # pragma: no cover - best effort
        try:
            lib = _lib.load()
            for ptr, _n in self._raw.values():
                lib.dcmoe_ipc_free(ptr)
        except Exception:  # noqa: BLE001
            pass


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
        if d.n_real % self.world != 0 or self.world > 8:
            raise ValueError(f"num_experts ({d.n_real}) must be divisible by ep_size ({self.world}) (core.py:505), ep_size <= 8")
        self.n_loc = d.n_real // self.world
        self.local_dims = replace(d, n_real=self.n_loc)
        self._w13 = self._w2 = None
        self.ws: Optional[EpWorkspace] = None
        self._peer = None
        self._imported = []
        self._flag = None
        self.kernels_per_step = 11  # router, plan, ep_plan, shared scales, dispatch, 4 x ffn gemm, combine x 2
        self.row_capacity = 0
        import os
        self.overlap = os.environ.get("DCMOE_EP_OVERLAP", "1") != "0"   # comm stream under the shared experts' GEMMs
        self.comm_ctas = int(os.environ.get("DCMOE_EP_COMM_CTAS", "148"))   # grid cap of dispatch / partial combine (0 = full)
        self.gemm_ctas = int(os.environ.get("DCMOE_EP_GEMM_CTAS", "0"))   # CTAs of the shared GEMMs that run under comm (0 = all SMs)
        # eighths of the shared experts' GEMM-1 that run under the dispatch, the rest then runs with GEMM-2 under the
        # combine gather; 0 = all of it under the dispatch.  Measured at 8 GPUs with 4/8: no gain (5.19 vs 4.86 ms per
        # step) -- the gather and the GEMMs compete for the same SMs and HBM, the gather just stretches from 0.70 to
        # 1.15 ms -- so the split stays off by default
        self.shared_split = int(os.environ.get("DCMOE_EP_SHARED_SPLIT", "0"))
        self._side = None
        self.comm_events = []        # (name, start, end) CUDA events of the comm-stream kernels, filled when a stage hook is set
        self._local_cfg = None
        self._partial = None
        # decode-sized calls (world * T <= 64 tokens, the same T on every rank): replicated routing of the gathered tokens,
        # local experts' rows computed by the weight-streaming kernels, combine over peer memory (see decode_forward)
        self.decode_mode = os.environ.get("DCMOE_EP_DECODE", "1") != "0"
        self._dws: Optional[EpWorkspace] = None
        self._dws_by_T = {}
        self._dy_raw = None               # (ptr, bytes) of the peer-visible y of the decode-sized path
        self._dpeer = None
        self._dflag = None

    # ------------------------------------------------------------------ setup
    def pack_local_weights(self):
        if self._w13 is not None:
            return
        m, d, ld = self.m, self.m.dims, self.local_dims
        p = m.gate.weight
        G = self.n_loc + 1
        w13 = torch.empty((G, 2 * d.dynamic_intermediate_size, d.hidden_size), dtype=p.dtype, device=p.device)
        w2 = torch.empty((G, d.hidden_size, d.dynamic_intermediate_size), dtype=p.dtype, device=p.device)
        routed, shared = m._expert_params()
        for l in range(self.n_loc):
            e = self.rank * self.n_loc + l
            ex = routed[e] if len(routed) == d.n_real else routed[l]   # full module or local-experts-only module
            ops.pack_expert(ex.gate_proj.weight.detach().contiguous(), ex.up_proj.weight.detach().contiguous(),
                            ex.down_proj.weight.detach().contiguous(), l, 0, ld, w13, w2)
        for i, ex in enumerate(shared):
            ops.pack_expert(ex.gate_proj.weight.detach().contiguous(), ex.up_proj.weight.detach().contiguous(),
                            ex.down_proj.weight.detach().contiguous(), self.n_loc, i, ld, w13, w2)
        self._w13, self._w2 = w13, w2

    def set_packed_local_weights(self, w13: torch.Tensor, w2: torch.Tensor):
        """Use packs built elsewhere (``checkpoint.load_dcmoe_ep``: this rank's experts read straight from a
        checkpoint) instead of packing from the wrapped module's parameters."""
        d, G = self.m.dims, self.n_loc + 1
        if tuple(w13.shape) != (G, 2 * d.dynamic_intermediate_size, d.hidden_size) or \
                tuple(w2.shape) != (G, d.hidden_size, d.dynamic_intermediate_size) or not (w13.is_cuda and w2.is_cuda):
            raise ValueError("packed local weights have the wrong shape for this rank's expert count")
        self._w13, self._w2 = w13.contiguous(), w2.contiguous()

    def default_row_capacity(self, T: int, T_global: int) -> int:
        t_pad = (T + 127) // 128 * 128
        return t_pad + self.n_loc * T_global + 128 * self.n_loc

    def ensure_workspace(self, T: int, dtype, device, ipc: bool, T_global: Optional[int] = None) -> EpWorkspace:
        cap = self.row_capacity or self.default_row_capacity(T, T_global if T_global is not None else T * self.world)
        if self.ws is None or self.ws.T != T or self.ws.dtype != dtype or self.ws.row_capacity != cap:
            self.ws = EpWorkspace(self.m.dims, dtype, T, device, cap, ipc)
            self._peer = None
        self.m.last_workspace = self.ws
        return self.ws

    def set_peers(self, x_packed: List[int], row_scale: List[int], y: List[int]):
        arr = lambda v: (ctypes.c_void_p * len(v))(*v)  # noqa: E731
        self._peer = (arr(x_packed), arr(row_scale), arr(y))

    def _exchange_handles(self):
        import torch.distributed as dist

        lib = _lib.load()
        mine = self.ws.export_handles()
        allh = [None] * self.world
        dist.all_gather_object(allh, mine, group=self.group)
        ptrs = {"x_packed": [], "row_scale": [], "y": []}
        for r in range(self.world):
            for name in ptrs:
                if r == self.rank:
                    ptrs[name].append(self.ws._raw[name][0])
                else:
                    p = ctypes.c_void_p()
                    buf = (ctypes.c_uint8 * 64).from_buffer_copy(allh[r][name])
                    _lib.check(lib.dcmoe_ipc_import(buf, ctypes.byref(p)), "dcmoe_ipc_import")
                    self._imported.append(p.value)
                    ptrs[name].append(p.value)
        self.set_peers(ptrs["x_packed"], ptrs["row_scale"], ptrs["y"])
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.ws.device)

    # ------------------------------------------------------------------ phases (all launch-only)
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
        self._all_counts = all_counts.contiguous()
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

    def phase_ffn(self, phase: int = 0, group_sel: int = 0, name: Optional[str] = None, max_ctas: int = 0,
                  shared_split: int = 0):
        """phase 0/1/2 = both / GEMM-1 / GEMM-2; group_sel 0/1/2/3 = all / shared-expert / routed row tiles / the shared
        tiles from the split point on; shared_split s (1..7): split point at s/8 of the shared tiles (group 1 stops there)."""
        lib = _lib.load()
        ws = self.ws
        if self._local_cfg is None:
            self._local_cfg = self.local_dims.c_config(ws.dtype)
        from .dcmoe import _DEFAULT_BF16_IMPL
        impl = self.m.ffn_impl if self.m.ffn_impl is not None else (_DEFAULT_BF16_IMPL if ws.dtype == torch.bfloat16 else 1)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.dcmoe_grouped_ffn(self._x.data_ptr(), ws.x_packed.data_ptr(), self._w13.data_ptr(),
                                         self._w2.data_ptr(), ws.row_scale.data_ptr(), ws.T, ws.row_capacity, self._local_cfg,
                                         ws.plan.data_ptr(), ws.h.data_ptr(), ws.y.data_ptr(), impl,
                                         phase | (group_sel << 4) | (max_ctas << 8) | (1 << 20) | (shared_split << 28), st),
                   "dcmoe_grouped_ffn")
        if name and self.m.stage_hook:
            self.m.stage_hook(name)

    def phase_combine(self, out: Optional[torch.Tensor], mode: int = 0, max_ctas: int = 0):
        lib = _lib.load()
        ws = self.ws
        _xp, _rs, y = self._peer
        if mode != 0 and self._partial is None:
            self._partial = torch.empty((max(ws.T, 1), self.m.dims.hidden_size), dtype=torch.float32, device=ws.device)
        _lib.check(lib.dcmoe_ep_combine(ws.y.data_ptr(), y, ws.slot_of.data_ptr(), ws.T, ws.cfg, self.world, mode,
                                        None if self._partial is None else self._partial.data_ptr(),
                                        None if out is None else out.data_ptr(), max_ctas,
                                        torch.cuda.current_stream().cuda_stream), "dcmoe_ep_combine")

    # ------------------------------------------------------------------ decode-sized calls
    def decode_applicable(self, T: int, dtype) -> bool:
        return (self.decode_mode and dtype == torch.bfloat16 and 0 < T * self.world <= 64 and self.m.use_front_small
                and self.m.ffn_impl in (None, 0, 3))

    def ensure_decode_workspace(self, T_total: int, device, ipc: bool) -> EpWorkspace:
        """Workspace of a decode-sized call on T_total gathered tokens (kept per token count).  Across processes the
        only peer-visible buffer is y; it is allocated ONCE, for 64 tokens, and every per-T workspace views it, so the
        cudaIpc handles are exchanged a single time however the batch size varies."""
        ws = self._dws_by_T.get(T_total)
        if ws is None:
            if len(self._dws_by_T) >= 8:
                self._dws_by_T.pop(next(iter(self._dws_by_T)))
            ws = EpWorkspace(self.m.dims, torch.bfloat16, T_total, device, 0, False)      # worst-case rows for T_total tokens
            ws.x_all = torch.empty((T_total, self.m.dims.hidden_size), dtype=torch.bfloat16, device=device)
            if ipc:
                H = self.m.dims.hidden_size
                if self._dy_raw is None:
                    lib = _lib.load()
                    sizes, _ = ops.query_sizes(self.m.dims, torch.bfloat16, 64, 0)
                    nbytes = int(sizes.row_capacity) * H * 2
                    p = ctypes.c_void_p()
                    with torch.cuda.device(device):
                        _lib.check(lib.dcmoe_ipc_alloc(nbytes, ctypes.byref(p)), "dcmoe_ipc_alloc")
                    self._dy_raw = (p.value, nbytes)
                assert ws.row_capacity * H * 2 <= self._dy_raw[1]
                ws.y = _wrap(self._dy_raw[0], ws.row_capacity * H * 2, torch.bfloat16, (ws.row_capacity, H), ws.device)
                ws._c_ws = None
            self._dws_by_T[T_total] = ws
        self._dws = ws
        return ws

    def set_decode_peers(self, y: List[int]):
        self._dpeer = (ctypes.c_void_p * len(y))(*y)

    def _exchange_decode_handles(self):
        import torch.distributed as dist

        lib = _lib.load()
        buf = (ctypes.c_uint8 * 64)()
        _lib.check(lib.dcmoe_ipc_export(self._dy_raw[0], buf), "dcmoe_ipc_export")
        allh = [None] * self.world
        dist.all_gather_object(allh, bytes(buf), group=self.group)
        ys = []
        for r in range(self.world):
            if r == self.rank:
                ys.append(self._dy_raw[0])
            else:
                p = ctypes.c_void_p()
                hb = (ctypes.c_uint8 * 64).from_buffer_copy(allh[r])
                _lib.check(lib.dcmoe_ipc_import(hb, ctypes.byref(p)), "dcmoe_ipc_import")
                self._imported.append(p.value)
                ys.append(p.value)
        self.set_decode_peers(ys)
        self._dflag = torch.zeros(1, dtype=torch.int32, device=self._dws.device)

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
        """Expert parallelism for decode-sized calls.  The dispatch / combine exchange of the large-T path costs eleven
        launches and three collectives (340 us per layer at T = 2); with a handful of tokens it is cheaper to replicate
        the tokens (one all-gather of world * T rows) and the routing, let every rank stream only ITS experts' weights,
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
        if self._dpeer is None:
            self._exchange_decode_handles()
        dist.all_gather_into_tensor(ws.x_all, x, group=self.group)     # also: every rank is done reading the previous y
        am_all = None
        if attention_mask is not None:
            am = attention_mask.reshape(-1).to(device=x.device, dtype=torch.int32).contiguous()
            am_all = torch.empty(T * self.world, dtype=torch.int32, device=x.device)
            dist.all_gather_into_tensor(am_all, am, group=self.group)
        self.decode_route_and_ffn(am_all)
        dist.all_reduce(self._dflag, group=self.group)                 # every owner's y rows are complete
        out = torch.empty((B, S, H), dtype=x.dtype, device=x.device)
        self.decode_combine(T, out)
        logits, top_k, mask, gw = (t[self.rank * T:(self.rank + 1) * T] for t in self._droute)
        self.m.last_workspace = ws
        return out, logits, top_k, mask, gw, ws.aux_loss.clone().reshape(())

    # ------------------------------------------------------------------ distributed forward
    @torch.no_grad()
    def __call__(self, hidden_states: torch.Tensor, attention_mask=None, aux_balance_weight=None, router_logits=None):
        import torch.distributed as dist

        if aux_balance_weight is not None:
            raise NotImplementedError("aux_balance_weight is training-only")
        B, S, H = hidden_states.shape
        T = B * S
        if router_logits is None and self.m.stage_hook is None and self.decode_applicable(T, hidden_states.dtype):
            out = self.decode_forward(hidden_states, attention_mask)
            if getattr(self.m, "avg_hidden_states_last", False):
                dist.all_reduce(out[0], op=dist.ReduceOp.SUM, group=self.group)
                out[0].div_(self.world)
            return out
        x = hidden_states.reshape(T, H)
        if not x.is_contiguous():
            x = x.contiguous()
        self.pack_local_weights()
        ws = self.ensure_workspace(T, x.dtype, x.device, ipc=True)
        if self._peer is None:
            self._exchange_handles()
        hook = self.m.stage_hook or (lambda _n: None)
        main = torch.cuda.current_stream()
        vec = self.phase_route(x, attention_mask, router_logits)
        all_counts = torch.empty((self.world, vec.numel()), dtype=torch.int32, device=x.device)
        dist.all_gather_into_tensor(all_counts, vec, group=self.group)          # also the "buffers are free" barrier
        hook("allgather_counts")
        self.phase_plan(all_counts)
        out = torch.empty((B, S, H), dtype=x.dtype, device=x.device)
        tcgen05 = (self.m.ffn_impl in (None, 0, 2)) and x.dtype == torch.bfloat16
        if not (self.overlap and tcgen05):
            self.phase_dispatch()
            hook("ep_dispatch")
            dist.all_reduce(self._flag, group=self.group)                        # every rank's rows have landed
            hook("barrier_dispatch")
            self.phase_ffn(1, 0, "ffn_gemm1")
            self.phase_ffn(2, 0, "ffn_gemm2")
            dist.all_reduce(self._flag, group=self.group)                        # every owner's y is complete
            hook("barrier_ffn")
            self.phase_combine(out.view(T, H), 0)
            hook("ep_combine")
        else:
            # comm stream: dispatch over NVLink + barrier, while the main stream runs the shared experts' GEMM-1
            if self._side is None:
                self._side = torch.cuda.Stream(x.device)
                self._ev = [torch.cuda.Event() for _ in range(4)]
                self._flag2 = torch.zeros(1, dtype=torch.int32, device=x.device)
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
                dist.all_reduce(self._flag2, group=self.group)                   # barrier 1 (on the comm stream)
                ev[1].record(side)
            ss = self.shared_split if self.m.ffn_impl in (None, 0) else 0
            self.phase_ffn(1, 1, "ffn_gemm1_shared", self.gemm_ctas, ss)         # overlaps the dispatch (first ss/8 of the tiles)
            main.wait_event(ev[1])
            hook("wait_dispatch")
            self.phase_ffn(1, 2, "ffn_gemm1_routed")
            self.phase_ffn(2, 2, "ffn_gemm2_routed")
            dist.all_reduce(self._flag, group=self.group)                        # barrier 2: routed y complete everywhere
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
            if ss:                                                               # rest of the shared GEMM-1 + GEMM-2 overlap
                self.phase_ffn(1, 3, "ffn_gemm1_shared_rest", self.gemm_ctas, ss)   # the combine gather
            self.phase_ffn(2, 1, "ffn_gemm2_shared", self.gemm_ctas)
            main.wait_event(ev[3])
            hook("wait_combine")
            self.phase_combine(out.view(T, H), 2)
            hook("ep_combine_final")
        if getattr(self.m, "avg_hidden_states_last", False):
            # core.py:355-356: all_reduce(final_hidden_states, AVG) over the expert-parallel group in eval mode
            dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
            out.div_(self.world)
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
        """The decode-sized expert-parallel path (ExpertParallelDCMoE.decode_forward) with the all-gather replaced by
        a torch.cat and the peers' y buffers by the virtual ranks' own.  Every x must have the same token count."""
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
    def forward(self, xs: Sequence[torch.Tensor], attention_masks=None):
        R = self.world
        shapes = [x.shape for x in xs]
        flat = [x.reshape(-1, x.shape[-1]).contiguous() for x in xs]
        T_global = sum(f.shape[0] for f in flat)
        for r, ep in enumerate(self.ranks):
            ep.pack_local_weights()
            ep.ensure_workspace(flat[r].shape[0], flat[r].dtype, flat[r].device, ipc=False, T_global=T_global)
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
                ep.phase_ffn(1, 1, shared_split=ep.shared_split)   # (part of the) shared GEMM-1 needs nothing from the dispatch
            for ep in self.ranks:
                ep.phase_dispatch()
            for ep in self.ranks:
                ep.phase_ffn(1, 2)
                ep.phase_ffn(2, 2)
            for ep in self.ranks:
                ep.phase_combine(None, 1)
            for ep in self.ranks:
                if ep.shared_split:
                    ep.phase_ffn(1, 3, shared_split=ep.shared_split)
                ep.phase_ffn(2, 1)
            outs = []
            for r, ep in enumerate(self.ranks):
                out = torch.empty_like(flat[r])
                ep.phase_combine(out, 2)
                logits, top_k, mask, gw = ep._route
                outs.append((out.view(shapes[r]), logits, top_k, mask, gw, ep._aux))
        self.all_counts = all_counts
        return outs
