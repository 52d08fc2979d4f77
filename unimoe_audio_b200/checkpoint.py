"""Checkpoint loading straight into the packed kernel layouts.

The reference loads weights through ``from_pretrained`` (utils/UniMoE_Audio_mod.py:79-83) into per-expert
``nn.Linear`` modules; training writes DeepSpeed MoE checkpoints with one file per (layer, expert)
(``layer_{L}_expert_{E}_mp_rank_00_model_states.pt`` next to ``mp_rank_00_model_states.pt``) which
``UniMoEV2-Preview/inference/deepspeed_ep_param_aggregation.py:16-49`` re-shards for expert parallelism: with
``g = source_ep_num // target_ep_size`` experts per rank, global expert ``e`` goes to rank ``e // g`` under the local
name ``e % g``.

``DCMoE`` already accepts the reference state dict (same keys), but that materialises every expert twice (module
parameters + packed W13 / W2).  The functions here read one expert at a time from the checkpoint and pack it
directly (``dcmoe_pack_expert``), so a layer only ever holds the packed copy, and an expert-parallel rank only
reads the experts it owns:

    src = SafetensorsSource("/path/to/checkpoint_dir")            # HF shards with the reference key names
    src = DeepSpeedSource("/path/to/global_step123")              # DeepSpeed MoE checkpoint directory
    layer = load_dcmoe(src, layer_id, config, torch.bfloat16, "cuda")                      # all experts
    layer, w13, w2 = load_dcmoe_ep(src, layer_id, config, torch.bfloat16, "cuda", rank, world)
    ep = ExpertParallelDCMoE(layer, group); ep.set_packed_local_weights(w13, w2)

Host-side logic only (file formats, key names, ownership); the packing itself is the CUDA kernel.
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass, replace
from typing import Dict, Iterable, List, Optional, Tuple

import torch

from . import ops
from .dcmoe import DCMoE, PostAttentionMoE

# reference key names (utils/UniMoE_Audio_core.py:220-222, deepspeed_ep_param_aggregation.py:18)
_ROUTED = re.compile(r"^(?P<prefix>.*layers\.(?P<layer>\d+)\.)mlp\.dynamic_real_moe\.deepspeed_moe\.experts\.deepspeed_experts\."
                     r"(?P<expert>\d+)\.(?P<proj>gate_proj|up_proj|down_proj)\.weight$")
_SHARED = re.compile(r"^(?P<prefix>.*layers\.(?P<layer>\d+)\.)mlp\.fixed_real_moe\.(?P<expert>\d+)\."
                     r"(?P<proj>gate_proj|up_proj|down_proj)\.weight$")
_GATE = re.compile(r"^(?P<prefix>.*layers\.(?P<layer>\d+)\.)mlp\.gate\.weight$")
_NORM = re.compile(r"^(?P<prefix>.*layers\.(?P<layer>\d+)\.)post_attention_layernorm\.weight$")
_EXPERT_FILE = re.compile(r"^layer_(\d+)_expert_(\d+)_mp_rank_00_model_states\.pt$")   # aggregation.py:17
_PROJS = ("gate_proj", "up_proj", "down_proj")


def expert_owner(expert_id: int, n_experts: int, ep_size: int) -> Tuple[int, int]:
    """(rank, local expert id) of a global routed expert -- deepspeed_ep_param_aggregation.py:21, :32-40."""
    if n_experts % ep_size != 0:
        raise ValueError(f"num_experts ({n_experts}) must be divisible by ep_size ({ep_size})")
    per_rank = n_experts // ep_size
    return expert_id // per_rank, expert_id % per_rank


class TensorSource:
    """A flat name -> tensor store read lazily, one tensor at a time."""

    def keys(self) -> Iterable[str]:
        raise NotImplementedError

    def get(self, key: str) -> torch.Tensor:
        raise NotImplementedError


class SafetensorsSource(TensorSource):
    """One or more ``.safetensors`` files (an HF checkpoint directory, or the files written by the reference's
    aggregation script) with the reference key names."""

    def __init__(self, path_or_files):
        from safetensors import safe_open

        self._open = safe_open
        if isinstance(path_or_files, (str, os.PathLike)):
            p = str(path_or_files)
            files = sorted(os.path.join(p, f) for f in os.listdir(p) if f.endswith(".safetensors")) if os.path.isdir(p) else [p]
        else:
            files = [str(f) for f in path_or_files]
        if not files:
            raise FileNotFoundError(f"no .safetensors files under {path_or_files}")
        self._where: Dict[str, str] = {}
        for f in files:
            with safe_open(f, framework="pt", device="cpu") as h:
                for k in h.keys():
                    self._where[k] = f

    def keys(self):
        return self._where.keys()

    def get(self, key: str) -> torch.Tensor:
        with self._open(self._where[key], framework="pt", device="cpu") as h:
            return h.get_tensor(key)


class DeepSpeedSource(TensorSource):
    """A DeepSpeed MoE checkpoint directory: ``mp_rank_00_model_states.pt`` (``["module"]`` = everything but the
    routed experts) plus ``layer_{L}_expert_{E}_mp_rank_00_model_states.pt`` per routed expert, whose keys carry the
    GLOBAL expert id (aggregation.py:17-19, :27-36).  Expert files are opened only when one of their tensors is asked
    for; the last one is cached."""

    def __init__(self, ckpt_dir: str):
        self.dir = str(ckpt_dir)
        main = os.path.join(self.dir, "mp_rank_00_model_states.pt")
        if not os.path.exists(main):
            raise FileNotFoundError(main)
        self._module = torch.load(main, map_location="cpu", weights_only=False)["module"]
        self._expert_files: Dict[Tuple[int, int], str] = {}
        for f in os.listdir(self.dir):
            mt = _EXPERT_FILE.match(f)
            if mt:
                self._expert_files[(int(mt.group(1)), int(mt.group(2)))] = os.path.join(self.dir, f)
        self._cache: Tuple[Optional[str], Optional[dict]] = (None, None)
        self._prefix = self._routed_prefix()

    def _routed_prefix(self) -> str:
        for k in self._module:
            g = _GATE.match(k)
            if g:
                return g.group("prefix")[: g.group("prefix").rfind("layers.")]
        return "model."

    def _expert_keys(self, layer: int, expert: int) -> List[str]:
        base = f"{self._prefix}layers.{layer}.mlp.dynamic_real_moe.deepspeed_moe.experts.deepspeed_experts.{expert}."
        return [base + p + ".weight" for p in _PROJS]

    def keys(self):
        out = list(self._module.keys())
        for (layer, expert) in sorted(self._expert_files):
            out.extend(self._expert_keys(layer, expert))
        return out

    def get(self, key: str) -> torch.Tensor:
        if key in self._module:
            return self._module[key]
        mt = _ROUTED.match(key)
        if not mt:
            raise KeyError(key)
        f = self._expert_files[(int(mt.group("layer")), int(mt.group("expert")))]
        if self._cache[0] != f:
            self._cache = (f, torch.load(f, map_location="cpu", weights_only=False))
        return self._cache[1][key]


@dataclass(frozen=True)
class LoadItem:
    """One expert to pack: the three checkpoint keys, and where it goes (weight group / part of ``dcmoe_pack_expert``)."""
    gate_proj: str
    up_proj: str
    down_proj: str
    group: int
    part: int


@dataclass(frozen=True)
class LayerPlan:
    layer_id: int
    prefix: str
    gate: str
    norm: Optional[str]
    items: Tuple[LoadItem, ...]
    n_local: int


def moe_layers(keys: Iterable[str]) -> Dict[int, str]:
    """layer id -> key prefix (up to and including ``layers.{L}.``) of every MoE layer in the checkpoint."""
    out: Dict[int, str] = {}
    for k in keys:
        g = _GATE.match(k)
        if g:
            out[int(g.group("layer"))] = g.group("prefix")
    return dict(sorted(out.items()))


def plan_layer_load(keys: Iterable[str], layer_id: int, n_real: int, n_fix: int, ep_rank: int = 0, ep_size: int = 1,
                    local_expert_ids: bool = False) -> LayerPlan:
    """Which checkpoint tensors rank ``ep_rank`` of ``ep_size`` packs for one layer, and into which weight group.

    Routed expert e is owned by rank ``e // (n_real // ep_size)`` and becomes local group ``e % (n_real // ep_size)``
    (aggregation.py:21, :32-40); the shared experts (group ``n_local``, parts 0..n_fix-1), the gate and the norm are
    replicated.  ``local_expert_ids``: the checkpoint was already re-sharded by the reference script for this rank
    (its routed keys carry local ids)."""
    keys = set(keys)
    layers = moe_layers(keys)
    if layer_id not in layers:
        raise KeyError(f"layer {layer_id} has no MoE block in this checkpoint (MoE layers: {list(layers)})")
    prefix = layers[layer_id]
    per_rank = n_real // ep_size
    if n_real % ep_size != 0 or not (0 <= ep_rank < ep_size):
        raise ValueError(f"bad expert-parallel layout: {n_real} experts, rank {ep_rank} of {ep_size}")

    def triple(fmt: str) -> Tuple[str, str, str]:
        names = tuple(fmt.format(proj=p) for p in _PROJS)
        for n in names:
            if n not in keys:
                raise KeyError(f"checkpoint is missing {n}")
        return names

    items: List[LoadItem] = []
    for e in range(n_real):
        rank, local = expert_owner(e, n_real, ep_size)
        if rank != ep_rank:
            continue
        src_id = local if local_expert_ids else e
        g, u, d = triple(f"{prefix}mlp.dynamic_real_moe.deepspeed_moe.experts.deepspeed_experts.{src_id}.{{proj}}.weight")
        items.append(LoadItem(g, u, d, local, 0))
    for i in range(n_fix):
        g, u, d = triple(f"{prefix}mlp.fixed_real_moe.{i}.{{proj}}.weight")
        items.append(LoadItem(g, u, d, per_rank, i))
    norm = f"{prefix}post_attention_layernorm.weight"
    return LayerPlan(layer_id, prefix, f"{prefix}mlp.gate.weight", norm if norm in keys else None, tuple(items), per_rank)


def _pack(source: TensorSource, plan: LayerPlan, dims: ops.LayerDims, dtype: torch.dtype, device) -> Tuple[torch.Tensor, torch.Tensor]:
    local = replace(dims, n_real=plan.n_local)
    G = plan.n_local + 1
    w13 = torch.empty((G, 2 * dims.dynamic_intermediate_size, dims.hidden_size), dtype=dtype, device=device)
    w2 = torch.empty((G, dims.hidden_size, dims.dynamic_intermediate_size), dtype=dtype, device=device)
    for it in plan.items:      # one expert on the device at a time
        g, u, d = (source.get(k).to(device=device, dtype=dtype).contiguous() for k in (it.gate_proj, it.up_proj, it.down_proj))
        want = dims.dynamic_intermediate_size if it.group < plan.n_local else dims.shared_intermediate_size
        if tuple(g.shape) != (want, dims.hidden_size) or tuple(u.shape) != (want, dims.hidden_size) or \
                tuple(d.shape) != (dims.hidden_size, want):
            raise ValueError(f"{it.gate_proj}: expert shapes {tuple(g.shape)}/{tuple(d.shape)} do not match the config")
        ops.pack_expert(g, u, d, it.group, it.part, local, w13, w2)
    return w13, w2


def _empty_layer(config, dtype, device) -> DCMoE:
    with torch.device("meta"):
        m = DCMoE(config)
    m = m.to(dtype)
    # only the gate is materialised; the per-expert Linear weights stay empty (the kernels read the packed copy)
    m.gate.weight = torch.nn.Parameter(torch.empty((m.num_experts, m.hidden_dim), dtype=dtype, device=device), requires_grad=False)
    routed, shared = m._expert_params()
    for ex in list(routed) + list(shared):
        for lin in (ex.gate_proj, ex.up_proj, ex.down_proj):
            lin.weight = torch.nn.Parameter(torch.empty(0, dtype=dtype, device=device), requires_grad=False)
    return m.eval()


def load_dcmoe(source: TensorSource, layer_id: int, config, dtype: torch.dtype = torch.bfloat16, device="cuda",
               with_norm: bool = False):
    """Build the MoE block of one decoder layer from a checkpoint, packing each expert as it is read.  With
    ``with_norm`` returns a :class:`PostAttentionMoE` (the layer's post_attention_layernorm is loaded too)."""
    device = torch.device(device)
    m = _empty_layer(config, dtype, device)
    plan = plan_layer_load(source.keys(), layer_id, m.dims.n_real, m.dims.n_fix)
    m.gate.weight.data.copy_(source.get(plan.gate).to(dtype))
    m._w13, m._w2 = _pack(source, plan, m.dims, dtype, device)
    m._reference_released = True
    if not with_norm:
        return m
    if plan.norm is None:
        raise KeyError(f"checkpoint has no {plan.prefix}post_attention_layernorm.weight")
    cfg = config if isinstance(config, dict) else config.__dict__
    blk = PostAttentionMoE({"rms_norm_eps": cfg.get("rms_norm_eps", 1e-6)}, mlp=m)
    blk.post_attention_layernorm.weight = torch.nn.Parameter(source.get(plan.norm).to(device=device, dtype=dtype),
                                                             requires_grad=False)
    return blk.eval()


def load_dcmoe_ep(source: TensorSource, layer_id: int, config, dtype: torch.dtype, device, ep_rank: int, ep_size: int,
                  local_expert_ids: bool = False):
    """Expert-parallel variant: returns ``(layer, w13_local, w2_local)`` where the packs hold this rank's routed
    experts (local group ids) plus the shared pair -- feed them to ``ExpertParallelDCMoE.set_packed_local_weights``.
    Only the owned experts are read from the checkpoint."""
    device = torch.device(device)
    m = _empty_layer(config, dtype, device)
    plan = plan_layer_load(source.keys(), layer_id, m.dims.n_real, m.dims.n_fix, ep_rank, ep_size, local_expert_ids)
    m.gate.weight.data.copy_(source.get(plan.gate).to(dtype))
    w13, w2 = _pack(source, plan, m.dims, dtype, device)
    m._reference_released = True
    return m, w13, w2
