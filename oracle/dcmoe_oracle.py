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


def route(logits, attention_mask=None, cfg: dict | None = None):
    c = dict(DEFAULT_CONFIG)
    c.update(cfg or {})
    n_dyn = c["mlp_dynamic_expert_num"] + c["mlp_dynamic_null_expert_num"]
    return route_oracle_c.route(logits, attention_mask, n_dyn=n_dyn, n_fix=c["mlp_fixed_expert_num"],
                                top_p=c["mlp_dynamic_top_p"], eps=c["router_jitter_noise"],
                                fixed_top_k=int(c.get("mlp_dynamic_top_k", 0) or 0))


@torch.no_grad()
def forward(hidden_states: torch.Tensor, weights: Dict[str, torch.Tensor], attention_mask: Optional[torch.Tensor] = None,
            cfg: dict | None = None, logits: Optional[torch.Tensor] = None, skip_ffn: bool = False) -> OracleOutput:
    """Eval-mode forward of ``UniMoEAudioSparseMoeBlock`` (core.py:236-358), token_drop=False.

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
        logits = F.linear(x, weights[GATE].to(D))                       # core.py:251
    am = None if attention_mask is None else attention_mask.reshape(-1)
    top_k, mask, gw, aux = route(logits, am, c)                         # core.py:255-332
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
