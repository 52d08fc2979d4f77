"""Host-buffer front end of the DCMoE layer: pinned host tensors in, pinned host tensors out.

``DCMoE.forward`` takes device tensors (that is what the decoder layer hands it, model.py:241).  A caller that
owns HOST activations (the end-to-end measurement of bench.py, or a CPU-resident pipeline stage) uses
``HostPipeline``: every step copies its input host->device, runs the layer and copies the whole 6-tuple
device->host, with the three phases of consecutive steps overlapped on three CUDA streams
(H2D of step i+1 and D2H of step i-1 run while step i computes; PCIe is full duplex).  Nothing is cached:
each step's bytes cross the bus inside the pipeline.
"""
from __future__ import annotations

from collections import deque
from typing import Callable, Deque, List, Optional, Tuple

import torch


class HostPipeline:
    def __init__(self, layer: Callable, depth: int = 3, device: Optional[torch.device] = None):
        # depth >= 3 keeps H2D(i+2), compute(i+1) and D2H(i) in flight together; with 2 the next H2D waits for a D2H
        self.layer = layer
        self.depth = depth
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.s_in = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.x_dev: List[Optional[torch.Tensor]] = [None] * depth
        self.out_host: List[Optional[List[torch.Tensor]]] = [None] * depth
        self.ev_h2d = [torch.cuda.Event() for _ in range(depth)]
        self.ev_compute = [torch.cuda.Event() for _ in range(depth)]
        self.ev_d2h = [torch.cuda.Event() for _ in range(depth)]
        self.pending: Deque[int] = deque()
        self.step = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._last_out = None

    def submit(self, x_host: torch.Tensor, attention_mask=None):
        """Enqueue one step.  ``x_host`` must be pinned.  Returns nothing; call ``result()`` in order."""
        if not x_host.is_pinned():
            raise ValueError("HostPipeline needs pinned host memory for asynchronous copies")
        slot = self.step % self.depth
        if len(self.pending) == self.depth:
            raise RuntimeError("pipeline full: call result() before submitting more steps")
        cur = torch.cuda.current_stream(self.device)
        if self.x_dev[slot] is None or self.x_dev[slot].shape != x_host.shape or self.x_dev[slot].dtype != x_host.dtype:
            self.x_dev[slot] = torch.empty(x_host.shape, dtype=x_host.dtype, device=self.device)
        # H2D on the copy-in stream, after the compute that last read this slot
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.ev_compute[slot])
            self.x_dev[slot].copy_(x_host, non_blocking=True)
            self.ev_h2d[slot].record(self.s_in)
        # compute on the caller's stream
        cur.wait_event(self.ev_h2d[slot])
        out = self.layer(self.x_dev[slot], attention_mask, None)
        self.ev_compute[slot].record(cur)
        # D2H on the copy-out stream
        if self.out_host[slot] is None or any(h.shape != t.shape for h, t in zip(self.out_host[slot], out)):
            self.out_host[slot] = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in out]
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_compute[slot])
            for h, t in zip(self.out_host[slot], out):
                t.record_stream(self.s_out)
                h.copy_(t, non_blocking=True)
            self.ev_d2h[slot].record(self.s_out)
        self.h2d_bytes = x_host.numel() * x_host.element_size()
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in out)
        self._last_out = out
        self.pending.append(slot)
        self.step += 1

    def result(self) -> Tuple[torch.Tensor, ...]:
        """Block until the oldest submitted step's outputs are in host memory and return them (the host
        tensors are reused ``depth`` steps later)."""
        slot = self.pending.popleft()
        self.ev_d2h[slot].synchronize()
        return tuple(self.out_host[slot])

    def copy_only_ms(self, x_host: torch.Tensor, steps: int) -> float:
        """Device time of ``steps`` rounds of one step's copies alone -- the H2D of ``x_host`` and the D2H of the last
        step's 6-tuple, on the two copy streams at once, no kernels: the host-side ceiling of the pipeline on this box
        (PCIe, host memory, NUMA placement), to be compared with the end-to-end step time."""
        if getattr(self, "_last_out", None) is None or self.pending:
            raise RuntimeError("copy_only_ms needs a drained pipeline that has run at least one step")
        torch.cuda.synchronize(self.device)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(self.s_in)
        ev[2].record(self.s_out)
        for _ in range(steps):
            with torch.cuda.stream(self.s_in):
                self.x_dev[0].copy_(x_host, non_blocking=True)
            with torch.cuda.stream(self.s_out):
                for h, t in zip(self.out_host[0], self._last_out):
                    h.copy_(t, non_blocking=True)
        ev[1].record(self.s_in)
        ev[3].record(self.s_out)
        torch.cuda.synchronize(self.device)
        return max(ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3]))

    def run(self, xs_host) -> List[Tuple[torch.Tensor, ...]]:
        outs = []
        for x in xs_host:
            if len(self.pending) == self.depth:
                outs.append(tuple(t.clone() for t in self.result()))
            self.submit(x)
        while self.pending:
            outs.append(tuple(t.clone() for t in self.result()))
        return outs


class GraphedDCMoE:
    """CUDA-graph replay of one DCMoE forward for a fixed token count (decode steps: T = 2N tokens per call,
    36 layers x up to 1000 steps -- reference model.py:1149-1203).  The forward is launch-only (four to six kernels, no
    host synchronisation, all data-dependent sizes in the device-side plan), so it captures as is; a replay costs
    one cudaGraphLaunch instead of six launches plus the Python around them.

        g = GraphedDCMoE(layer, batch, seq, dtype)      # captures on the current device
        out = g(hidden_states)                           # same 6-tuple; tensors are reused by the next call
    """

    def __init__(self, layer: Callable, batch: int, seq: int, dtype: torch.dtype = torch.bfloat16,
                 hidden: int = 2048, device: Optional[torch.device] = None, with_mask: bool = False):
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.x = torch.zeros((batch, seq, hidden), dtype=dtype, device=self.device)
        self.mask = torch.ones((batch, seq), dtype=torch.int32, device=self.device) if with_mask else None
        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(2):                       # warm-up: packs weights, allocates the workspace
                layer(self.x, self.mask, None)
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = layer(self.x, self.mask, None)
        # the captured kernels point into the layer's workspace: keep it alive as long as the graph
        self.workspace = getattr(layer, "last_workspace", None)

    @torch.no_grad()
    def __call__(self, hidden_states: torch.Tensor, attention_mask: Optional[torch.Tensor] = None):
        self.x.copy_(hidden_states, non_blocking=True)
        if self.mask is not None:
            if attention_mask is None:
                self.mask.fill_(1)
            else:
                self.mask.copy_(attention_mask.reshape(self.mask.shape).to(torch.int32), non_blocking=True)
        elif attention_mask is not None:
            raise ValueError("this graph was captured without an attention mask (with_mask=False)")
        self.graph.replay()
        return self.out
