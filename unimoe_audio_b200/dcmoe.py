"""Drop-in DCMoE layer: the reference's ``UniMoEAudioSparseMoeBlock`` on hand-written sm_100a CUDA.

Mirrors reference utils/UniMoE_Audio_core.py:196-358 at the boundary SURVEY.md section 8(b) describes:

  * constructor: ``DCMoE(config)`` with ``config`` the (duck-typed) ``Qwen2_5_VLMoETextConfig`` built from
    utils/config.json["text_config"]; the same attributes are read (core.py:204-234, :24, :42, :501);
  * parameters / state-dict keys: ``gate.weight``, ``fixed_real_moe.{i}.{gate,up,down}_proj.weight``,
    ``dynamic_real_moe.deepspeed_moe.experts.deepspeed_experts.{e}.{gate,up,down}_proj.weight``
    so ``from_pretrained`` / ``load_state_dict`` of real checkpoints work unchanged;
  * call: ``forward(hidden_states, attention_mask, aux_balance_weight)`` -> the 6-tuple
    ``(final_hidden_states, full_router_logits, dynamic_top_k, expert_mask, global_weight, aux_loss)``
    with the reference's dtypes (x.dtype, x.dtype, int64, int32, x.dtype, float32 0-dim) (core.py:358).

Swap-in (no edit to UniMoE_Audio_mod.py / UniMoE_Audio_model.py): either rebind
``utils.UniMoE_Audio_model.UniMoEAudioSparseMoeBlock = DCMoE`` before the model is built, or after loading
``for l in model.language_model.layers: l.mlp = DCMoE.from_reference(l.mlp)`` -- see INTEGRATION.md.

The forward is launch-only (router -> plan -> permute -> grouped FFN -> combine, six kernels, no host
synchronisation) and has NO CPU fallback: inputs must be CUDA tensors and the CUDA extension must load.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from .ops import LayerDims, Workspace


def _cfg_get(config, name, default=None):
    if isinstance(config, dict):
        return config.get(name, default)
    return getattr(config, name, default)


class _ExpertMLP(nn.Module):
    """Parameter holder with the reference's names (AudioSharedExpertMLP core.py:16-31 /
    AudioDynamicExpertMLP core.py:34-49).  Compute happens in the grouped kernels, not here."""

    def __init__(self, hidden_size: int, intermediate_size: int):
        super().__init__()
        self.hidden_size, self.intermediate_size = hidden_size, intermediate_size
        self.gate_proj = nn.Linear(hidden_size, intermediate_size, bias=False)
        self.up_proj = nn.Linear(hidden_size, intermediate_size, bias=False)
        self.down_proj = nn.Linear(intermediate_size, hidden_size, bias=False)

    def forward(self, *_a, **_k):  # pragma: no cover - never used as a module
        raise RuntimeError("expert MLPs are executed by the grouped sm_100a kernels, not individually")


class _Experts(nn.Module):          # AudioExperts, core.py:392-416
    def __init__(self, hidden_size, intermediate_size, num_local_experts):
        super().__init__()
        self.deepspeed_experts = nn.ModuleList([_ExpertMLP(hidden_size, intermediate_size) for _ in range(num_local_experts)])
        self.num_local_experts = num_local_experts


class _MOELayer(nn.Module):         # AudioMOELayer, core.py:419-493
    def __init__(self, experts, ep_size, num_local_experts):
        super().__init__()
        self.experts = experts
        self.ep_group = None
        self.ep_size = ep_size
        self.num_local_experts = num_local_experts

    def _set_ep_group(self, ep_group):
        self.ep_group = ep_group


class _MoE(nn.Module):              # UniMoEAudioMoE, core.py:496-523
    def __init__(self, hidden_size, intermediate_size, num_experts, ep_size):
        super().__init__()
        if num_experts % ep_size != 0:
            raise ValueError(f"num_experts ({num_experts}) must be divisible by ep_size ({ep_size})")
        self.ep_size = ep_size
        self.num_experts = num_experts
        self.num_local_experts = num_experts // ep_size
        self.deepspeed_moe = _MOELayer(_Experts(hidden_size, intermediate_size, self.num_local_experts), ep_size,
                                       self.num_local_experts)


_WORKSPACES: Dict[tuple, Workspace] = {}


def _workspace_bytes(ws: Workspace) -> int:
    return sum(t.numel() * t.element_size() for t in (ws.plan, ws.h, ws.x_packed, ws.y, ws.row_scale, ws.slot_of, ws.row_token))


def get_workspace(dims: LayerDims, dtype: torch.dtype, T: int, device, row_capacity: int = 0) -> Workspace:
    """Workspaces are shared between layers (the 36 decoder layers run back to back on one stream) and kept per token
    count: least recently used first out, at most 16 of them and 8 GiB in total (decode-sized ones are ~16 MB, a
    16,384-token one ~2 GB)."""
    key = (dims, dtype, T, torch.device(device), row_capacity)
    ws = _WORKSPACES.pop(key, None)
    if ws is None:
        ws = Workspace(dims, dtype, T, device, row_capacity)
        need = _workspace_bytes(ws)
        while _WORKSPACES and (len(_WORKSPACES) >= 16 or
                               need + sum(_workspace_bytes(w) for w in _WORKSPACES.values()) > (8 << 30)):
            _WORKSPACES.pop(next(iter(_WORKSPACES)))
    _WORKSPACES[key] = ws          # (re)insert at the most-recently-used end
    return ws


class DCMoE(nn.Module):
    """B200-native replacement of ``UniMoEAudioSparseMoeBlock`` (core.py:196)."""

    def __init__(self, config, ffn_impl: Optional[int] = None):
        super().__init__()
        g = lambda n, d=None: _cfg_get(config, n, d)  # noqa: E731
        # ---- the attributes the reference reads (core.py:204-234) ----
        self.hidden_dim = g("hidden_size")
        self.mlp_dynamic_real_expert_num = g("mlp_dynamic_expert_num")
        self.mlp_dynamic_null_expert_num = g("mlp_dynamic_null_expert_num")
        self.mlp_dynamic_expert_num = self.mlp_dynamic_real_expert_num + self.mlp_dynamic_null_expert_num
        self.mlp_dynamic_top_p = g("mlp_dynamic_top_p")
        self.mlp_dynamic_top_k = g("mlp_dynamic_top_k")
        self.mlp_fixed_expert_num = g("mlp_fixed_expert_num")
        self.num_experts = self.mlp_dynamic_expert_num + self.mlp_fixed_expert_num
        self.ignore_differentiable_router = g("ignore_differentiable_router", True)
        self.router_jitter_noise = g("router_jitter_noise")
        self.input_jitter_noise = g("input_jitter_noise", 0.0)
        self.min_capacity = g("min_capacity", 8)
        self.capacity_factor = g("capacity_factor", 1.0)
        self.token_drop = g("token_drop", False)
        self.drop_policy = g("drop_policy", "probs")
        self.avg_hidden_states_last = g("avg_hidden_states_last", False)
        self.drop_token_num_print = g("drop_token_num_print", False)
        self.fp32_gate = g("fp32_gate", False)
        self.ep_size = g("ep_size", 1)
        if self.mlp_dynamic_top_p == 0 and not (1 <= int(self.mlp_dynamic_top_k or 0)):
            raise ValueError("mlp_dynamic_top_p == 0 (fixed top-k routing, core.py:256-257) needs mlp_dynamic_top_k >= 1")
        if self.token_drop and self.drop_policy == "position":
            # core.py:321-323 applies `cumsum(expert_mask) - 1 < capacity` to the shared experts' all-ones columns too:
            # most tokens past the capacity lose all eleven columns and the reference returns NaN global weights and
            # hidden states for them (tests/golden/drop_position_nan.npz) -- there is no result to be compatible with
            raise NotImplementedError("token_drop with drop_policy='position' (core.py:321-323) yields NaN outputs in the "
                                      "reference for tokens past the capacity; only drop_policy='probs' is implemented")
        if self.avg_hidden_states_last and self.ep_size == 1:
            raise NotImplementedError("avg_hidden_states_last=True (core.py:355-356) averages over the expert-parallel "
                                      "group and needs ep_size > 1")
        if g("hidden_act", "silu") != "silu":
            raise NotImplementedError("only hidden_act='silu' (utils/config.json) is implemented")
        self.dims = LayerDims(
            hidden_size=self.hidden_dim, n_real=self.mlp_dynamic_real_expert_num, n_null=self.mlp_dynamic_null_expert_num,
            n_fix=self.mlp_fixed_expert_num, dynamic_intermediate_size=g("dynamic_intermediate_size"),
            shared_intermediate_size=g("shared_intermediate_size"), top_p=float(self.mlp_dynamic_top_p),
            jitter_eps=float(self.router_jitter_noise),
            fixed_top_k=int(self.mlp_dynamic_top_k or 0) if self.mlp_dynamic_top_p == 0 else 0)
        # ---- parameters under the reference's names (core.py:220-222) ----
        self.gate = nn.Linear(self.hidden_dim, self.num_experts, bias=False)
        self.fixed_real_moe = nn.ModuleList(
            [_ExpertMLP(self.hidden_dim, self.dims.shared_intermediate_size) for _ in range(self.mlp_fixed_expert_num)])
        self.dynamic_real_moe = _MoE(self.hidden_dim, self.dims.dynamic_intermediate_size,
                                     self.mlp_dynamic_real_expert_num, self.ep_size)
        # ---- packed weights for the grouped kernels (built lazily from the parameters) ----
        self._w13: Optional[torch.Tensor] = None
        self._w2: Optional[torch.Tensor] = None
        self._packed_key = None
        self.ffn_impl = ffn_impl          # None: tcgen05 for bf16, CUDA-core for fp32
        self.row_capacity = 0             # rows of the FFN workspace; 0 = from row_capacity_factor
        # routed rows per token the workspace is sized for; None = worst case (n_real: every token to every routed
        # expert -- ~2 GB at 16,384 tokens, 8 GB at 65,536).  Top-P 0.7 routing averages 3.6; with a factor below n_real
        # an overflow (rows dropped) is detected and raised by the next forward / check_overflow()
        env_f = __import__("os").environ.get("DCMOE_ROW_CAPACITY_FACTOR")
        self.row_capacity_factor: Optional[float] = float(env_f) if env_f else None
        self.last_workspace: Optional[Workspace] = None
        self.stage_hook = None            # optional callable(stage_name) invoked between kernel launches (bench)
        self._reference_released = False
        self._ep = None                   # ExpertParallelDCMoE, built on the first forward when ep_size > 1
        self.use_front_small = True       # T <= 64 (bf16): fused router + plan + permute launch

    # ------------------------------------------------------------------ weights
    @classmethod
    def from_reference(cls, block: nn.Module, config=None, **kw) -> "DCMoE":
        """Build from an instantiated reference block (or any module with the same state dict)."""
        if config is None:
            exp0 = block.dynamic_real_moe.deepspeed_moe.experts.deepspeed_experts[0]
            config = dict(
                hidden_size=block.hidden_dim, mlp_dynamic_expert_num=block.mlp_dynamic_real_expert_num,
                mlp_dynamic_null_expert_num=block.mlp_dynamic_null_expert_num, mlp_dynamic_top_p=block.mlp_dynamic_top_p,
                mlp_dynamic_top_k=block.mlp_dynamic_top_k, mlp_fixed_expert_num=block.mlp_fixed_expert_num,
                ignore_differentiable_router=block.ignore_differentiable_router,
                router_jitter_noise=block.router_jitter_noise, input_jitter_noise=block.input_jitter_noise,
                min_capacity=block.min_capacity, capacity_factor=block.capacity_factor, token_drop=block.token_drop,
                drop_policy=block.drop_policy, avg_hidden_states_last=block.avg_hidden_states_last,
                drop_token_num_print=block.drop_token_num_print, fp32_gate=block.fp32_gate,
                ep_size=getattr(block.dynamic_real_moe, "ep_size", 1),
                dynamic_intermediate_size=exp0.gate_proj.weight.shape[0],
                shared_intermediate_size=block.fixed_real_moe[0].gate_proj.weight.shape[0], hidden_act="silu")
        new = cls(config, **kw)
        ref_param = next(block.parameters())
        new.to(device=ref_param.device, dtype=ref_param.dtype)
        new.load_state_dict(block.state_dict())
        return new.eval()

    def _expert_params(self):
        routed = self.dynamic_real_moe.deepspeed_moe.experts.deepspeed_experts
        return routed, self.fixed_real_moe

    def pack_weights(self, force: bool = False):
        """(Re)build W13 [n_real+1, 2*I_d, H] (gate/up rows interleaved in blocks of 64) and
        W2 [n_real+1, H, I_d]; group n_real is the shared-expert pack (2 x I_s = I_d)."""
        if self._w13 is not None and self._reference_released:
            return
        p = self.gate.weight
        routed, shared = self._expert_params()
        key = (p.device, p.dtype, tuple(w._version for m in list(routed) + list(shared)
                                        for w in (m.gate_proj.weight, m.up_proj.weight, m.down_proj.weight)))
        if not force and self._w13 is not None and self._packed_key == key:
            return
        if not p.is_cuda:
            raise RuntimeError("DCMoE needs its parameters on a CUDA device (no CPU fallback)")
        d = self.dims
        G = d.n_real + 1
        w13 = torch.empty((G, 2 * d.dynamic_intermediate_size, d.hidden_size), dtype=p.dtype, device=p.device)
        w2 = torch.empty((G, d.hidden_size, d.dynamic_intermediate_size), dtype=p.dtype, device=p.device)
        for e, m in enumerate(routed):
            ops.pack_expert(m.gate_proj.weight.detach().contiguous(), m.up_proj.weight.detach().contiguous(),
                            m.down_proj.weight.detach().contiguous(), e, 0, d, w13, w2)
        for i, m in enumerate(shared):
            ops.pack_expert(m.gate_proj.weight.detach().contiguous(), m.up_proj.weight.detach().contiguous(),
                            m.down_proj.weight.detach().contiguous(), d.n_real, i, d, w13, w2)
        self._w13, self._w2, self._packed_key = w13, w2, key

    def release_reference_weights(self):
        """Free the per-expert gate/up/down parameters after packing (halves the weight memory of a layer: the
        kernels only read the packed W13 / W2).  The state dict can no longer be saved from this module."""
        self.pack_weights()
        routed, shared = self._expert_params()
        for m in list(routed) + list(shared):
            for lin in (m.gate_proj, m.up_proj, m.down_proj):
                lin.weight.data = torch.empty(0, dtype=lin.weight.dtype, device=lin.weight.device)
        self._reference_released = True

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, hidden_states: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                aux_balance_weight: Optional[torch.Tensor] = None, router_logits: Optional[torch.Tensor] = None,
                residual: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """core.py:236-358.  In ``train()`` mode this is the FORWARD of the training step only (values, no autograd
        graph): the fp32 gate and the input jitter of core.py:240-249 run, the mixer takes its eval branch because
        ``ignore_differentiable_router`` is set (core.py:272; utils/config.json:71)."""
        if self.training and not self.ignore_differentiable_router:
            raise NotImplementedError("training-mode forward with ignore_differentiable_router=False (gumbel sampling and "
                                      "the differentiable routing function, core.py:111-135) is not implemented")
        if self.token_drop and self.drop_policy != "probs":
            raise ValueError(f"Invalid drop_policy: {self.drop_policy}")          # core.py:325
        if hidden_states.dim() != 3 or hidden_states.shape[-1] != self.hidden_dim:
            raise ValueError(f"hidden_states must be [batch, seq, {self.hidden_dim}]")
        if not hidden_states.is_cuda:
            raise RuntimeError("DCMoE.forward needs CUDA tensors: there is no CPU fallback")
        dt = hidden_states.dtype
        if dt != self.gate.weight.dtype:
            raise TypeError(f"hidden_states dtype {dt} != parameter dtype {self.gate.weight.dtype}")
        if hidden_states.device != self.gate.weight.device:
            raise RuntimeError(f"hidden_states on {hidden_states.device} but the layer's parameters on {self.gate.weight.device}")
        # launches go to the CURRENT device on the C side: make the input's device current (as PyTorch ops do)
        with ops.on_device(hidden_states.device):
            return self._forward(hidden_states, attention_mask, aux_balance_weight, router_logits, residual)

    def _forward(self, hidden_states, attention_mask, aux_balance_weight, router_logits, residual):
        dt = hidden_states.dtype
        if self.ep_size != 1:
            # expert parallelism configured the reference's way (config.ep_size, core.py:505-520): this rank holds
            # num_experts / ep_size routed experts and the forward runs ep.ExpertParallelDCMoE over the group set by
            # deepspeed_moe._set_ep_group (default: the world group)
            if router_logits is not None or residual is not None:
                raise NotImplementedError("router_logits / residual are single-GPU extras")
            if self._ep is None:
                import torch.distributed as dist

                from .ep import ExpertParallelDCMoE
                if not dist.is_initialized():
                    raise RuntimeError("ep_size > 1 needs an initialised torch.distributed process group (NCCL)")
                group = self.dynamic_real_moe.deepspeed_moe.ep_group or dist.group.WORLD
                if dist.get_world_size(group) != self.ep_size:
                    raise ValueError(f"ep_size ({self.ep_size}) != size of the expert-parallel group "
                                     f"({dist.get_world_size(group)})")
                self._ep = ExpertParallelDCMoE(self, group)
            return self._ep(hidden_states, attention_mask, aux_balance_weight)
        self.pack_weights()
        return self._forward_local(hidden_states, attention_mask, aux_balance_weight, router_logits, residual,
                                   self._w13, self._w2)

    def _forward_local(self, hidden_states, attention_mask, aux_balance_weight, router_logits, residual, w13, w2,
                       before_ffn=None):
        """The whole layer on this GPU with the packed weights ``w13`` / ``w2`` (all n_real + 1 groups).  ``before_ffn``
        (weight-gather expert parallelism) is called after the permute launch and before the first GEMM launch: it
        makes the stream wait for the staged weights."""
        dt = hidden_states.dtype
        B, S, H = hidden_states.shape
        T = B * S
        x = hidden_states.reshape(T, H)
        if not x.is_contiguous():
            x = x.contiguous()
        ws = get_workspace(self.dims, dt, T, x.device, self.effective_row_capacity(T))
        self.last_workspace = ws
        if ws.reduced:
            ws.raise_if_overflowed()
        wg = self.gate.weight.detach()
        if not wg.is_contiguous():
            wg = wg.contiguous()
        hook = self.stage_hook or (lambda _name: None)
        out = torch.empty((B, S, H), dtype=dt, device=x.device)
        res = None
        if residual is not None:      # extension: fuse the decoder layer's residual add (model.py:242)
            if residual.shape != hidden_states.shape or residual.dtype != dt:
                raise ValueError("residual must match hidden_states in shape and dtype")
            res = residual.reshape(T, H)
            if not res.is_contiguous():
                res = res.contiguous()
        impl = self.ffn_impl if self.ffn_impl is not None else (0 if dt == torch.bfloat16 else 1)
        # ---- training-mode forward (core.py:240-249): fp32 gate, input jitter (torch's own RNG, as the reference) ----
        gate_fp32, x_gate = False, None
        if self.training:
            eps_in = float(self.input_jitter_noise or 0.0)
            if self.fp32_gate and dt != torch.float32:
                gate_fp32 = True
                if eps_in > 0:      # the jitter multiplies the float COPY: only the gate sees it (core.py:241-244)
                    x_gate = x.float()
                    x_gate *= torch.empty_like(x_gate).uniform_(1.0 - eps_in, 1.0 + eps_in)
            elif eps_in > 0:        # in place on the caller's tensor, as core.py:244 does
                hidden_states *= torch.empty_like(hidden_states).uniform_(1.0 - eps_in, 1.0 + eps_in)
                x = hidden_states.reshape(T, H)
                if not x.is_contiguous():
                    x = x.contiguous()
        extended = self.token_drop or aux_balance_weight is not None or gate_fp32
        if (not extended and self.stage_hook is None and before_ffn is None and router_logits is None and T > 0 and
                (self.use_front_small or T > 64 or dt != torch.bfloat16)):
            # the common case: the whole layer in one host call (dcmoe_forward) -- same launches as below
            logits, top_k, mask, gw, aux = ops.forward(x, wg, w13, w2, ws, out, attention_mask, res, impl)
            if ws.reduced and not torch.cuda.is_current_stream_capturing():
                ws.note_overflow_async()
            if self.mlp_dynamic_top_p == 0:
                top_k = top_k.to(torch.int32)
            return out, logits, top_k, mask, gw, aux
        hook("start")
        small = (dt == torch.bfloat16 and 0 < T <= 64 and router_logits is None and self.use_front_small and not extended)
        aux_fixed = None
        if small:     # decode-sized call: router + plan + permute in one single-CTA launch
            logits, top_k, mask, gw = ops.front_small(x, wg, ws, attention_mask=attention_mask)
            hook("front_small")
        else:
            if x_gate is not None and router_logits is None:
                # fp32 gate on the jittered float copy: fp32 x / W_g -> fp32 logits; routing then runs on those logits
                router_logits = ops.router(x_gate, wg.float(), ws, attention_mask=attention_mask, all_fp32=True)[0]
            logits, top_k, mask, gw = ops.router(x, wg, ws, logits_in=router_logits, attention_mask=attention_mask,
                                                 fp32_gate=gate_fp32)
            hook("router")
            ops.plan(ws)
            hook("plan")
            if aux_balance_weight is not None or gate_fp32:
                # core.py:380-385 (weighted means), on the mask BEFORE any token drop (:293); with the fp32 gate the plain
                # means are taken here too, in fp32 (the plan kernel rounds the mean probability to the layer dtype)
                aux_fixed = ops.aux_weighted(logits, mask, aux_balance_weight, ws)
            if self.token_drop and T > 0:
                # core.py:302-329: capacity -> per-expert top-`capacity` tokens by logit -> AND, renormalise, new global
                # weights; the aux loss above is the one of the un-dropped mask
                if aux_fixed is None:
                    aux_fixed = ws.aux_loss.clone().reshape(())
                cap = ops.expert_capacity(self.dims, T, self.capacity_factor, self.min_capacity)
                keep = ops.drop_select(logits, mask, cap, ws)
                before = mask[:, : self.dims.n_dyn].sum() if self.drop_token_num_print else None
                logits, top_k, mask, gw = ops.router(None, None, ws, logits_in=logits, attention_mask=attention_mask,
                                                     keep=keep, fp32_gate=gate_fp32)
                ops.plan(ws)
                hook("token_drop")
                if before is not None and int(__import__("os").environ.get("RANK", "0")) == 0:     # core.py:316-319 (syncs, as there)
                    ori = int(before.item())
                    print(f"drop {ori - int(mask[:, : self.dims.n_dyn].sum().item())} tokens from total {ori} tokens")
        if T > 0:
            if not small:
                ops.permute(x, mask, gw, ws)
                hook("permute")
            if before_ffn is not None:
                before_ffn()
                hook("wait_weights")
            if self.stage_hook is None:
                ops.grouped_ffn(x, w13, w2, ws, impl, phase=0)     # both GEMMs, one call
            else:
                ops.grouped_ffn(x, w13, w2, ws, impl, phase=1)
                hook("ffn_gemm1")
                ops.grouped_ffn(x, w13, w2, ws, impl, phase=2)
                hook("ffn_gemm2")
            if aux_fixed is not None:
                aux = aux_fixed
                ops.combine(ws, out, res)
            else:
                aux = torch.empty((), dtype=torch.float32, device=x.device)
                ops.combine(ws, out, res, aux_out=aux)      # the 4-byte aux copy rides in the combine launch
            hook("combine")
        else:
            aux = aux_fixed if aux_fixed is not None else ws.aux_loss.clone().reshape(())
        if ws.reduced and not torch.cuda.is_current_stream_capturing():
            ws.note_overflow_async()
        if self.mlp_dynamic_top_p == 0:
            top_k = top_k.to(torch.int32)        # the reference builds it with torch.full(..., dtype=torch.int) (core.py:257)
        return out, logits, top_k, mask, gw, aux

    def effective_row_capacity(self, T: int) -> int:
        """Rows of the FFN workspace for a T-token call: ``row_capacity`` if set, else from ``row_capacity_factor``
        (0 = the worst case, which can never overflow)."""
        if self.row_capacity:
            return int(self.row_capacity)
        f = self.row_capacity_factor
        if f is None or f >= self.dims.n_real or T <= 64:
            return 0
        t_pad = (T + 127) // 128 * 128
        routed = int(-(-f * T // 1))
        return t_pad + (routed + 127) // 128 * 128 + 128 * self.dims.n_real

    def check_overflow(self):
        """Block until the last forward's overflow flag is on the host and raise if rows were dropped."""
        if self.last_workspace is not None and self.last_workspace.reduced:
            self.last_workspace.raise_if_overflowed(block=True)


class PostAttentionMoE(nn.Module):
    """Host mirror of the second half of the reference decoder layer (utils/UniMoE_Audio_model.py:239-242):

        residual = h;  h = self.post_attention_layernorm(h);  h, *router = self.mlp(h, mask, aux_w);  h = residual + h

    with the same sub-module names (``post_attention_layernorm.weight``, ``mlp.*``), so a decoder layer's state dict
    loads into it unchanged.  The RMSNorm is one HBM pass (``dcmoe_rmsnorm``) and the residual add is fused into the
    combine pass (``dcmoe_combine(residual=...)``): no separate elementwise kernels run around the MoE block.
    Returns the 6-tuple of the MoE block with ``hidden_states`` already holding ``residual + mlp(norm(h))``.
    """

    class _Norm(nn.Module):
        def __init__(self, hidden_size: int, eps: float):
            super().__init__()
            self.weight = nn.Parameter(torch.ones(hidden_size))
            self.variance_epsilon = eps

    def __init__(self, config, mlp: Optional["DCMoE"] = None):
        super().__init__()
        cfg = config if isinstance(config, dict) else config.__dict__
        self.mlp = mlp if mlp is not None else DCMoE(config)
        self.post_attention_layernorm = PostAttentionMoE._Norm(self.mlp.hidden_dim, float(cfg.get("rms_norm_eps", 1e-6)))

    @classmethod
    def from_reference_layer(cls, layer: nn.Module) -> "PostAttentionMoE":
        """Build from a reference decoder layer (model.py:196-247): takes its ``mlp`` and ``post_attention_layernorm``."""
        mlp = layer.mlp if isinstance(layer.mlp, DCMoE) else DCMoE.from_reference(layer.mlp)
        eps = float(layer.post_attention_layernorm.variance_epsilon)
        m = cls({"rms_norm_eps": eps}, mlp=mlp)
        w = layer.post_attention_layernorm.weight.detach()
        m.post_attention_layernorm.weight = nn.Parameter(w.clone().to(mlp.gate.weight.device), requires_grad=False)
        return m.eval()

    @torch.no_grad()
    def forward(self, hidden_states: torch.Tensor, padding_token_mask: Optional[torch.Tensor] = None,
                aux_balance_weight: Optional[torch.Tensor] = None):
        if not hidden_states.is_cuda:
            raise RuntimeError("PostAttentionMoE.forward needs CUDA tensors: there is no CPU fallback")
        x = hidden_states if hidden_states.is_contiguous() else hidden_states.contiguous()
        w = self.post_attention_layernorm.weight.detach()
        if w.dtype != x.dtype:
            raise TypeError(f"hidden_states dtype {x.dtype} != norm weight dtype {w.dtype}")
        with ops.on_device(x.device):
            normed = ops.rmsnorm(x, w, self.post_attention_layernorm.variance_epsilon, self.mlp.dims)
            return self.mlp(normed, padding_token_mask, aux_balance_weight, residual=x)


# The reference's class name, so `utils.UniMoE_Audio_model.UniMoEAudioSparseMoeBlock = ...` reads naturally.
UniMoEAudioSparseMoeBlock = DCMoE
