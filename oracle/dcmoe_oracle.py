"""CPU restatement of the whole DCMoE layer forward (test infrastructure only).

ORACLE -- NOT A PRODUCT PATH.  Only ``tests/``, ``tools/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module.  The product
(``unimoe_audio_b200``) never imports anything from ``oracle/`` and has no CPU fallback.

Parity status: PINNED against the unmodified reference block run in the build container
(fixtures in ``tests/golden/`` produced by ``tools/make_golden.py``; checked by
``tests/test_oracle_golden.py``).

Follows, function by function (paths relative to the reference tree):
  gate projection ........................ utils/UniMoE_Audio_core.py:251
  Top-P count / mixer / weights / aux ..... oracle/route_oracle.c  (core.py:157-167, :94-154,
                                            :259-291, :178-193, :361-389)
  dispatch (compress_matrix) .............. utils/UniMoE_Audio_utils.py:436-485, core.py:455-462
  routed experts (SwiGLU FFN) ............. core.py:406-416, :48-49
  combine (decompress + einsum) ........... utils/UniMoE_Audio_utils.py:488-523, core.py:486-488
  shared experts .......................... core.py:344-351, :30-31
  decoder-layer glue (rmsnorm, residual) .. utils/UniMoE_Audio_model.py:239-242; Qwen2RMSNorm.forward of the
                                            reference's pinned dependency transformers (model.py:54, :207), pinned by
                                            tests/golden/glue_*.npz (tools/make_golden_glue.py)

Differences from the reference *implementation* (not its results): rows are gathered per expert
with ``index_select`` in ascending token order (the canonical stable permutation of SURVEY.md
section 8a-7) instead of materialising ``[T, 8, H]`` temporaries, which lets the oracle run
config-2-sized inputs in seconds.  The FFN is row independent, so results are unchanged.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import route_oracle_c

DEFAULT_CONFIG = dict(
    hidden_size=2048,
    mlp_dynamic_expert_num=8,
    mlp_dynamic_null_expert_num=1,
    mlp_dynamic_top_p=0.7,
    mlp_dynamic_top_k=0.0,
    mlp_fixed_expert_num=2,
    dynamic_intermediate_size=2752,
    shared_intermediate_size=1376,
    router_jitter_noise=0.01,
    hidden_act="silu",
)

GATE = "gate.weight"
SHARED = "fixed_real_moe.{e}.{proj}.weight"
ROUTED = "dynamic_real_moe.deepspeed_moe.experts.deepspeed_experts.{e}.{proj}.weight"


def make_weights(cfg: dict | None = None, seed: int = 0, std: float = 0.02, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """N(0, std^2) weights under the reference's state-dict keys, generated in the
    same order as ``oracle.ref_loader.build_reference_block`` so both sides hold equal tensors."""
    c = dict(DEFAULT_CONFIG)
    c.update(cfg or {})
    H, Id, Is = c["hidden_size"], c["dynamic_intermediate_size"], c["shared_intermediate_size"]
    E = c["mlp_dynamic_expert_num"] + c["mlp_dynamic_null_expert_num"] + c["mlp_fixed_expert_num"]
    shapes = {GATE: (E, H)}
    for e in range(c["mlp_fixed_expert_num"]):
        shapes[SHARED.format(e=e, proj="gate_proj")] = (Is, H)
        shapes[SHARED.format(e=e, proj="up_proj")] = (Is, H)
        shapes[SHARED.format(e=e, proj="down_proj")] = (H, Is)
    for e in range(c["mlp_dynamic_expert_num"]):
        shapes[ROUTED.format(e=e, proj="gate_proj")] = (Id, H)
        shapes[ROUTED.format(e=e, proj="up_proj")] = (Id, H)
        shapes[ROUTED.format(e=e, proj="down_proj")] = (H, Id)
    gen = torch.Generator().manual_seed(seed)
    out = {}
    for name in sorted(shapes):
        out[name] = (torch.randn(shapes[name], generator=gen, dtype=torch.float32) * std).to(dtype)
    return out


@dataclass
class OracleOutput:
    final_hidden_states: torch.Tensor
    full_router_logits: torch.Tensor
    dynamic_top_k: torch.Tensor
    expert_mask: torch.Tensor
    global_weight: torch.Tensor
    aux_loss: torch.Tensor
    counts: torch.Tensor          # per routed expert token counts  (core.py:455 before .max())
    permutation: list             # canonical stable permutation: per expert, ascending token ids

    def as_tuple(self):
        return (self.final_hidden_states, self.full_router_logits, self.dynamic_top_k, self.expert_mask,
                self.global_weight, self.aux_loss)


def _ffn(x, wg, wu, wd):
    return F.linear(F.silu(F.linear(x, wg)) * F.linear(x, wu), wd)


def canonical_permutation(expert_mask: torch.Tensor, n_real: int):
    """Stable dispatch order: for each routed expert, token indices in ascending order
    (= ``argsort(mask.float(), dim=0, descending=True, stable=True)[:count]``; the reference's
    utils.py:460 uses the non-stable argsort, whose within-expert order is implementation
    defined -- SURVEY.md 8a-7)."""
    return [torch.nonzero(expert_mask[:, e], as_tuple=True)[0] for e in range(n_real)]


def expert_capacity(num_tokens: int, n_dyn: int, capacity_factor: float, min_capacity: int) -> int:
    """core.py:170-175 (`_audio_expert_capacity`) + the clamp of core.py:306-308: the quotient is a Python float, the
    product and the ceil run on a 0-dim float32 tensor."""
    import numpy as np

    cap = int(np.ceil(np.float32(num_tokens / n_dyn) * np.float32(capacity_factor)))
    if cap < int(min_capacity):
        cap = int(min_capacity)
    return min(cap, int(num_tokens))


def drop_keep_mask(logits: torch.Tensor, expert_mask: torch.Tensor, n_dyn: int, capacity: int) -> torch.Tensor:
    """Capacity mask of drop_policy == "probs" (core.py:309-313): per dynamic column, the `capacity` tokens with the
    largest logit among the tokens that selected the expert (unselected tokens enter torch.topk with finfo.min and are
    cleared again by the AND of :314).  torch.topk(sorted=False) leaves the choice among TIED boundary logits to the
    implementation; the pinned rule here (and in the CUDA kernel) is: ties go to the lower token index.  Returns
    uint8 [T, E] (shared columns 1)."""
    T, E = logits.shape
    keep = torch.ones((T, E), dtype=torch.uint8)
    lf = logits.float()
    for e in range(n_dyn):
        sel = expert_mask[:, e] != 0
        col = torch.zeros(T, dtype=torch.uint8)
        idx = torch.nonzero(sel, as_tuple=True)[0]
        if idx.numel() <= capacity:
            col[idx] = 1
        else:
            order = torch.sort(lf[idx, e], descending=True, stable=True).indices     # stable: lower token first on ties
            col[idx[order[:capacity]]] = 1
        keep[:, e] = col
    return keep


def route(logits, attention_mask=None, cfg: dict | None = None, aux_balance_weight=None):
    c = dict(DEFAULT_CONFIG)
    c.update(cfg or {})
    n_dyn = c["mlp_dynamic_expert_num"] + c["mlp_dynamic_null_expert_num"]
    kw = dict(n_dyn=n_dyn, n_fix=c["mlp_fixed_expert_num"], top_p=c["mlp_dynamic_top_p"], eps=c["router_jitter_noise"],
              fixed_top_k=int(c.get("mlp_dynamic_top_k", 0) or 0),
              aux_weight=None if aux_balance_weight is None else aux_balance_weight.reshape(-1))
    out = route_oracle_c.route(logits, attention_mask, **kw)
    if not c.get("token_drop", False):
        return out
    policy = c.get("drop_policy", "probs")
    if policy != "probs":
        # "position" (core.py:321-323) multiplies the shared experts' all-ones columns by the capacity test too: every
        # token with index >= capacity loses all columns and the reference returns NaN weights for it
        # (tests/golden/drop_position_nan.npz) -- nothing to restate
        raise ValueError(f"drop_policy {policy!r}: only 'probs' is restated")
    cap = expert_capacity(logits.shape[0], n_dyn, c.get("capacity_factor", 1.0), c.get("min_capacity", 8))
    keep = drop_keep_mask(logits, out[1], n_dyn, cap)                    # from the PRE-drop mask (core.py:309)
    top_k, mask, gw, _aux = route_oracle_c.route(logits, attention_mask, keep=keep, **kw)
    return top_k, mask, gw, out[3]                                      # the aux loss is computed before the drop (:293)


@torch.no_grad()
def forward(hidden_states: torch.Tensor, weights: Dict[str, torch.Tensor], attention_mask: Optional[torch.Tensor] = None,
            cfg: dict | None = None, logits: Optional[torch.Tensor] = None, skip_ffn: bool = False,
            aux_balance_weight: Optional[torch.Tensor] = None, fp32_gate: bool = False) -> OracleOutput:
    """Forward of ``UniMoEAudioSparseMoeBlock`` (core.py:236-358): eval mode, or -- ``fp32_gate=True`` -- the
    training-mode forward with the fp32 gate of core.py:240-249 (input_jitter_noise = 0, ignore_differentiable_router:
    no random branch runs).  ``cfg["token_drop"]`` selects the capacity branch (core.py:302-329, "probs" policy).

    ``logits`` may be supplied to pin the router input ("bit-exact given identical router logits").
    """
    c = dict(DEFAULT_CONFIG)
    c.update(cfg or {})
    B, S, H = hidden_states.shape
    D = hidden_states.dtype
    n_real = c["mlp_dynamic_expert_num"]
    n_dyn = n_real + c["mlp_dynamic_null_expert_num"]
    n_fix = c["mlp_fixed_expert_num"]
    x = hidden_states.reshape(-1, H)
    T = x.shape[0]
    if logits is None:
        if fp32_gate:
            logits = F.linear(x.float(), weights[GATE].float())         # core.py:241, :249
        else:
            logits = F.linear(x, weights[GATE].to(D))                   # core.py:251
    am = None if attention_mask is None else attention_mask.reshape(-1)
    top_k, mask, gw, aux = route(logits, am, c, aux_balance_weight)     # core.py:255-332
    gw = gw.to(D)                                                       # core.py:339
    perm = canonical_permutation(mask, n_real)
    counts = mask[:, :n_real].sum(0)
    final = torch.zeros((T, H), dtype=D)
    if not skip_ffn:
        rw = gw[:, :n_real] * mask[:, :n_real].to(D)                    # core.py:447
        comb = torch.zeros((T, H), dtype=torch.float32)                 # einsum "se,sem->sm": fp32 accumulate
        for e in range(n_real):
            idx = perm[e]
            if idx.numel() == 0:
                continue
            y = _ffn(x.index_select(0, idx),
                     weights[ROUTED.format(e=e, proj="gate_proj")].to(D),
                     weights[ROUTED.format(e=e, proj="up_proj")].to(D),
                     weights[ROUTED.format(e=e, proj="down_proj")].to(D))
            comb.index_add_(0, idx, rw[idx, e, None].float() * y.float())
        final = final + comb.to(D)                                      # core.py:342
        for e in range(n_fix):                                          # core.py:344-351
            y = _ffn(x, weights[SHARED.format(e=e, proj="gate_proj")].to(D),
                     weights[SHARED.format(e=e, proj="up_proj")].to(D),
                     weights[SHARED.format(e=e, proj="down_proj")].to(D))
            final = final + y * gw[:, n_dyn + e, None]
    return OracleOutput(final.reshape(B, S, H), logits, top_k, mask, gw, aux, counts, perm)


def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """post_attention_layernorm (reference model.py:240): transformers Qwen2RMSNorm.forward restated --
    fp32 mean of squares, rsqrt, cast to the layer dtype, multiply by the weight in the layer dtype."""
    dt = x.dtype
    xf = x.to(torch.float32)
    var = xf.pow(2).mean(-1, keepdim=True)
    n = (xf * torch.rsqrt(var + eps)).to(dt)
    return weight * n


def glue_forward(hidden_states: torch.Tensor, norm_weight: torch.Tensor, weights: Dict[str, torch.Tensor],
                 attention_mask: Optional[torch.Tensor] = None, eps: float = 1e-6, cfg: dict | None = None,
                 logits: Optional[torch.Tensor] = None):
    """Second half of the decoder layer (reference model.py:239-242): residual + mlp(rmsnorm(h)).
    Returns (hidden_states_out, OracleOutput of the MoE call)."""
    n = rmsnorm(hidden_states, norm_weight, eps)
    o = forward(n, weights, attention_mask, cfg, logits=logits)
    return hidden_states + o.final_hidden_states.reshape(hidden_states.shape), o
