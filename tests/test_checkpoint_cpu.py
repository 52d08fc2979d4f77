"""Host logic of unimoe_audio_b200.checkpoint (file formats, key names, expert ownership) on synthetic checkpoints,
checked against the restatement of the reference's re-sharding script (oracle/ep_reshard_oracle.py)."""
import os

import pytest
import torch

from oracle import ep_reshard_oracle as RO
from unimoe_audio_b200 import checkpoint as C

H, ID, IS, NR, NF, LAYERS = 256, 128, 64, 8, 2, (0, 3)
PRE = "model.layers.{L}."


def _layer_tensors(L, seed):
    g = torch.Generator().manual_seed(seed + L)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    p = PRE.format(L=L)
    t = {p + "mlp.gate.weight": r(NR + 1 + NF, H), p + "post_attention_layernorm.weight": r(H),
         p + "self_attn.q_proj.weight": r(8, 8)}
    for i in range(NF):
        for proj, shape in (("gate_proj", (IS, H)), ("up_proj", (IS, H)), ("down_proj", (H, IS))):
            t[p + f"mlp.fixed_real_moe.{i}.{proj}.weight"] = r(*shape)
    e = {}
    for x in range(NR):
        for proj, shape in (("gate_proj", (ID, H)), ("up_proj", (ID, H)), ("down_proj", (H, ID))):
            e[(x, p + f"mlp.dynamic_real_moe.deepspeed_moe.experts.deepspeed_experts.{x}.{proj}.weight")] = r(*shape)
    return t, e


@pytest.fixture(scope="module")
def ckpts(tmp_path_factory):
    from safetensors.torch import save_file
    root = tmp_path_factory.mktemp("ckpt")
    ds = root / "deepspeed"; hf = root / "hf"
    ds.mkdir(); hf.mkdir()
    module, flat, files = {}, {}, {}
    for L in LAYERS:
        t, e = _layer_tensors(L, 100)
        module.update(t); flat.update(t)
        for x in range(NR):
            d = {k: v for (xx, k), v in e.items() if xx == x}
            fn = f"layer_{L}_expert_{x}_mp_rank_00_model_states.pt"
            torch.save(d, ds / fn)
            files[fn] = list(d)
            flat.update(d)
    torch.save({"module": module}, ds / "mp_rank_00_model_states.pt")
    ks = sorted(flat)
    save_file({k: flat[k].contiguous() for k in ks[: len(ks) // 2]}, str(hf / "model-00001-of-00002.safetensors"))
    save_file({k: flat[k].contiguous() for k in ks[len(ks) // 2:]}, str(hf / "model-00002-of-00002.safetensors"))
    return dict(ds=str(ds), hf=str(hf), flat=flat, module=list(module), files=files)


def test_sources_expose_the_same_tensors(ckpts):
    a, b = C.SafetensorsSource(ckpts["hf"]), C.DeepSpeedSource(ckpts["ds"])
    assert set(a.keys()) == set(b.keys()) == set(ckpts["flat"])
    for k in list(ckpts["flat"])[::7]:
        assert torch.equal(a.get(k), ckpts["flat"][k]) and torch.equal(b.get(k), ckpts["flat"][k])
    assert C.moe_layers(a.keys()) == {L: PRE.format(L=L) for L in LAYERS}


@pytest.mark.parametrize("ep_size", [1, 2, 4, 8])
def test_ownership_matches_the_reference_resharding_rule(ckpts, ep_size):
    src = C.DeepSpeedSource(ckpts["ds"])
    target = RO.aggregation_names(ckpts["module"], ckpts["files"], source_ep_num=NR, target_ep_size=ep_size)
    for rank in range(ep_size):
        for L in LAYERS:
            plan = C.plan_layer_load(src.keys(), L, NR, NF, rank, ep_size)
            assert plan.n_local == NR // ep_size
            routed = [it for it in plan.items if it.group < plan.n_local]
            assert [it.group for it in routed] == list(range(NR // ep_size))
            for it in routed:      # the tensor the reference script files under local id `group` on this rank
                for proj, key in (("gate_proj", it.gate_proj), ("up_proj", it.up_proj), ("down_proj", it.down_proj)):
                    tname = f"model.layers.{L}.mlp.dynamic_real_moe.deepspeed_moe.experts.deepspeed_experts.{it.group}.{proj}.weight"
                    assert target[rank][tname] == key
                    assert C.expert_owner(int(key.split("deepspeed_experts.")[1].split(".")[0]), NR, ep_size) == (rank, it.group)
            shared = [it for it in plan.items if it.group == plan.n_local]
            assert [it.part for it in shared] == list(range(NF))
            assert plan.gate in target[rank] and plan.norm in target[rank]


def test_resharded_checkpoint_with_local_ids(ckpts):
    """A checkpoint already written by the reference script for rank 1 of 2 carries LOCAL expert ids."""
    keys = [k for k in ckpts["module"]]
    for L in LAYERS:
        for l in range(NR // 2):
            for proj in ("gate_proj", "up_proj", "down_proj"):
                keys.append(f"model.layers.{L}.mlp.dynamic_real_moe.deepspeed_moe.experts.deepspeed_experts.{l}.{proj}.weight")
    plan = C.plan_layer_load(keys, 3, NR, NF, ep_rank=1, ep_size=2, local_expert_ids=True)
    assert [it.gate_proj.split("deepspeed_experts.")[1][0] for it in plan.items[:4]] == ["0", "1", "2", "3"]


def test_errors(ckpts):
    src = C.SafetensorsSource(ckpts["hf"])
    with pytest.raises(KeyError):
        C.plan_layer_load(src.keys(), 1, NR, NF)                 # layer 1 has no MoE block
    with pytest.raises(ValueError):
        C.plan_layer_load(src.keys(), 0, NR, NF, 0, 3)           # 8 experts over 3 ranks
    with pytest.raises(KeyError):
        C.plan_layer_load([k for k in src.keys() if "deepspeed_experts.5.up_proj" not in k], 0, NR, NF)
    with pytest.raises(FileNotFoundError):
        C.DeepSpeedSource(os.path.dirname(ckpts["ds"]))
